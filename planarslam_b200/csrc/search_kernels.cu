// Host side of the projection-guided searches (kernels: search_kernels.cuh).
#include "search_kernels.cuh"

namespace pslam {

// ---------------------------------------------------------------------------------------------------------
struct SearchBuffers {
    void* blob = nullptr; size_t cap = 0;
    uint32_t* d_cand = nullptr; int32_t* d_cand_n = nullptr; size_t cap_pts = 0;
    int32_t *d_cell_start = nullptr, *d_items = nullptr, *d_matches = nullptr, *d_scalar = nullptr, *d_hist_idx = nullptr;
    int8_t* d_hist_bin = nullptr; uint8_t* d_in_view = nullptr;
};

template <typename T>
static T* carve(uint8_t*& p, size_t n) { T* r = reinterpret_cast<T*>(p); p += (n * sizeof(T) + 255) / 256 * 256; return r; }

static int ensure(pslam_ctx* c, SearchBuffers& B, size_t npts) {
    if (!B.d_cell_start) {
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_cell_start, (SG_CELLS + 1) * sizeof(int32_t)));
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_items, SEARCH_MAX_KP * sizeof(int32_t)));
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_matches, SEARCH_MAX_KP * sizeof(int32_t)));
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_scalar, 4 * sizeof(int32_t)));
    }
    if (npts > B.cap_pts) {
        cudaFree(B.d_cand); cudaFree(B.d_cand_n); cudaFree(B.d_hist_idx); cudaFree(B.d_hist_bin); cudaFree(B.d_in_view);
        const size_t n = npts * 3 / 2 + 64;
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_cand, n * SEARCH_CAND_CAP * sizeof(uint32_t)));
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_cand_n, n * sizeof(int32_t)));
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_hist_idx, n * sizeof(int32_t)));
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_hist_bin, n));
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_in_view, n));
        B.cap_pts = n;
    }
    return PSLAM_OK;
}

// uploads the frame / map arrays into one staging blob; fills the device views
static int upload(pslam_ctx* c, SearchBuffers& B, const pslam_frame_view* f, const pslam_map_points* m, const pslam_last_frame* l, SearchFrameDev& F,
                  SearchMapDev& M, SearchLastDev& Ld) {
    const size_t nf = f->n, nm = m->n, nl = l ? l->n : 0;
    const size_t need = 4096 + nf * (28 + 4 + 32) + nm * (12 + 12 + 4 + 4 + 32 + 1 + 1) + nl * (28 + 4 + 1) + 256 * 16;
    if (need > B.cap) { cudaFree(B.blob); B.blob = nullptr; B.cap = 0; PSLAM_CUDA(c, cudaMalloc(&B.blob, need * 3 / 2)); B.cap = need * 3 / 2; }
    uint8_t* p = (uint8_t*)B.blob;
    cudaStream_t st = c->stream;
#define UP(dst, src, type, count) { type* d_ = carve<type>(p, (count) ? (count) : 1); if (count) PSLAM_CUDA(c, cudaMemcpyAsync(d_, src, (size_t)(count) * sizeof(type), cudaMemcpyHostToDevice, st)); dst = d_; }
    F.n = (int)nf;
    UP(F.keys_un, f->keys_un, pslam_keypoint, nf); UP(F.u_right, f->u_right, float, nf); UP(F.desc, f->desc, uint8_t, nf * 32);
    std::memcpy(F.Tcw, f->Tcw, sizeof F.Tcw); F.Tcw_dev = nullptr; F.n_dev = nullptr;
    F.fx = f->fx; F.fy = f->fy; F.cx = f->cx; F.cy = f->cy; F.bf = f->bf; F.min_x = f->min_x; F.max_x = f->max_x; F.min_y = f->min_y; F.max_y = f->max_y;
    F.n_levels = f->n_levels; F.log_scale_factor = f->log_scale_factor;
    for (int i = 0; i < f->n_levels && i < PSLAM_MAX_LEVELS; ++i) F.scale[i] = f->scale_factors[i];
    F.inv_w = (float)SG_COLS / (f->max_x - f->min_x); F.inv_h = (float)SG_ROWS / (f->max_y - f->min_y);
    M.n = (int)nm;
    UP(M.pos, m->pos, float, nm * 3); UP(M.normal, m->normal, float, nm * 3); UP(M.max_distance, m->max_distance, float, nm);
    UP(M.min_distance, m->min_distance, float, nm); UP(M.desc, m->desc, uint8_t, nm * 32); UP(M.skip, m->skip, uint8_t, nm); UP(M.has_obs, m->has_obs, uint8_t, nm);
    if (l) {
        Ld.n = (int)nl;
        UP(Ld.keys, l->keys, pslam_keypoint, nl); UP(Ld.map_point, l->map_point, int32_t, nl); UP(Ld.outlier, l->outlier, uint8_t, nl);
        std::memcpy(Ld.Tcw, l->Tcw, sizeof Ld.Tcw); Ld.Tcw_dev = nullptr; Ld.n_dev = nullptr;
    }
#undef UP
    return PSLAM_OK;
}

static int validate(pslam_ctx* c, const pslam_frame_view* f, const pslam_map_points* m) {
    if (!f || !m || f->n < 0 || m->n < 0) return set_error(c, PSLAM_E_INVALID, "null view");
    if (f->n > SEARCH_MAX_KP || f->n > 65535) return set_error(c, PSLAM_E_INVALID, "more than 4096 keypoints in the frame view");
    if (f->n_levels < 1 || f->n_levels > PSLAM_MAX_LEVELS || !(f->max_x > f->min_x) || !(f->max_y > f->min_y)) return set_error(c, PSLAM_E_INVALID, "bad frame view");
    return PSLAM_OK;
}

}  // namespace pslam

using namespace pslam;

extern "C" {

int pslam_search_by_projection_map(pslam_ctx* c, const pslam_frame_view* f, const pslam_map_points* m, float th, float nnratio, int32_t* matches_io,
                                   uint8_t* in_view) {
    if (!c) return PSLAM_E_INVALID;
    int rc = validate(c, f, m);
    if (rc != PSLAM_OK) return rc;
    if (!matches_io) return set_error(c, PSLAM_E_INVALID, "null matches");
    if (f->n == 0 || m->n == 0) { if (in_view && m->n) std::memset(in_view, 0, m->n); return 0; }
    PSLAM_CUDA(c, cudaSetDevice(c->cfg.device));
    if (!c->search) c->search = new SearchBuffers();
    SearchBuffers& B = *c->search;
    if ((rc = ensure(c, B, (size_t)m->n)) != PSLAM_OK) return rc;
    SearchFrameDev F; SearchMapDev M; SearchLastDev Ld;
    if ((rc = upload(c, B, f, m, nullptr, F, M, Ld)) != PSLAM_OK) return rc;
    cudaStream_t st = c->stream;
    PSLAM_CUDA(c, cudaMemcpyAsync(B.d_matches, matches_io, f->n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    PSLAM_CUDA(c, cudaMemsetAsync(B.d_scalar, 0, 4 * sizeof(int32_t), st));
    PSLAM_LAUNCH(c, "search_grid", k_search_grid<<<1, 1024, 0, st>>>(F, B.d_cell_start, B.d_items));
    PSLAM_LAUNCH(c, "search_candidates_map", k_candidates_map<<<(M.n + 7) / 8, 256, 0, st>>>(F, M, th, B.d_cell_start, B.d_items, B.d_cand, B.d_cand_n,
                 B.d_in_view, B.d_scalar + 1));
    PSLAM_LAUNCH(c, "search_resolve_map", k_resolve_map<<<1, 32, 0, st>>>(M, nnratio, B.d_cand, B.d_cand_n, B.d_matches, B.d_scalar));
    int32_t sc[2] = {0, 0};
    PSLAM_CUDA(c, cudaMemcpyAsync(matches_io, B.d_matches, f->n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (in_view) PSLAM_CUDA(c, cudaMemcpyAsync(in_view, B.d_in_view, m->n, cudaMemcpyDeviceToHost, st));
    PSLAM_CUDA(c, cudaMemcpyAsync(sc, B.d_scalar, sizeof sc, cudaMemcpyDeviceToHost, st));
    PSLAM_CUDA(c, cudaStreamSynchronize(st));
    if (sc[1]) return set_error(c, PSLAM_E_CAPACITY, "more than 128 candidates in one search window");
    return sc[0];
}

int pslam_search_by_projection_last(pslam_ctx* c, const pslam_frame_view* cur, const pslam_last_frame* last, const pslam_map_points* m, float th,
                                    int mono, int check_orientation, int32_t* matches_io) {
    if (!c) return PSLAM_E_INVALID;
    int rc = validate(c, cur, m);
    if (rc != PSLAM_OK) return rc;
    if (!last || last->n < 0 || !matches_io) return set_error(c, PSLAM_E_INVALID, "null last-frame view or matches");
    if (cur->n == 0 || last->n == 0) return 0;
    PSLAM_CUDA(c, cudaSetDevice(c->cfg.device));
    if (!c->search) c->search = new SearchBuffers();
    SearchBuffers& B = *c->search;
    if ((rc = ensure(c, B, (size_t)last->n)) != PSLAM_OK) return rc;
    SearchFrameDev F; SearchMapDev M; SearchLastDev Ld;
    if ((rc = upload(c, B, cur, m, last, F, M, Ld)) != PSLAM_OK) return rc;
    cudaStream_t st = c->stream;
    PSLAM_CUDA(c, cudaMemcpyAsync(B.d_matches, matches_io, cur->n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    PSLAM_CUDA(c, cudaMemsetAsync(B.d_scalar, 0, 4 * sizeof(int32_t), st));
    PSLAM_LAUNCH(c, "search_grid", k_search_grid<<<1, 1024, 0, st>>>(F, B.d_cell_start, B.d_items));
    PSLAM_LAUNCH(c, "search_candidates_last", k_candidates_last<<<(Ld.n + 7) / 8, 256, 0, st>>>(F, Ld, M, th, mono, B.d_cell_start, B.d_items, B.d_cand,
                 B.d_cand_n, B.d_scalar + 1));
    PSLAM_LAUNCH(c, "search_resolve_last", k_resolve_last<<<1, 32, 0, st>>>(F, Ld, M, check_orientation, B.d_cand, B.d_cand_n, B.d_matches, B.d_scalar,
                 B.d_hist_idx, B.d_hist_bin));
    int32_t sc[2] = {0, 0};
    PSLAM_CUDA(c, cudaMemcpyAsync(matches_io, B.d_matches, cur->n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PSLAM_CUDA(c, cudaMemcpyAsync(sc, B.d_scalar, sizeof sc, cudaMemcpyDeviceToHost, st));
    PSLAM_CUDA(c, cudaStreamSynchronize(st));
    if (sc[1]) return set_error(c, PSLAM_E_CAPACITY, "more than 128 candidates in one search window");
    return sc[0];
}

}  // extern "C"

namespace pslam {
void search_free(pslam_ctx* c) {
    if (!c->search) return;
    SearchBuffers& B = *c->search;
    cudaFree(B.blob); cudaFree(B.d_cand); cudaFree(B.d_cand_n); cudaFree(B.d_cell_start); cudaFree(B.d_items); cudaFree(B.d_matches);
    cudaFree(B.d_scalar); cudaFree(B.d_hist_idx); cudaFree(B.d_hist_bin); cudaFree(B.d_in_view);
    delete c->search;
    c->search = nullptr;
}
}  // namespace pslam
