// TEST INFRASTRUCTURE ONLY — C entry points (ctypes) onto the oracle line-segment detector.
#include "lsd.h"
#include <cstring>
using namespace oracle;
extern "C" {
// segs: float[cap][4]; wpn: double[cap][3] (width, precision, log-NFA); returns the number of segments found
// rect_enum: 0 published LSD rectangle iterator, 1 cv2 4.13's enumeration (lsd.cc rect_nfa)
int orc_lsd_detect_enum(const uint8_t* img, int w, int h, int stride, int refine, int rect_enum, float* segs, double* wpn, int cap) {
    std::vector<LsdSegment> out;
    lsd_detect(Img8{img, w, h, stride}, refine, out, rect_enum);
    const int n = (int)out.size();
    for (int i = 0; i < n && i < cap; ++i) {
        segs[4 * i] = out[i].x1; segs[4 * i + 1] = out[i].y1; segs[4 * i + 2] = out[i].x2; segs[4 * i + 3] = out[i].y2;
        wpn[3 * i] = out[i].width; wpn[3 * i + 1] = out[i].prec; wpn[3 * i + 2] = out[i].nfa;
    }
    return n;
}
int orc_lsd_detect(const uint8_t* img, int w, int h, int stride, int refine, float* segs, double* wpn, int cap) {
    return orc_lsd_detect_enum(img, w, h, stride, refine, 0, segs, wpn, cap);
}
int orc_lsd_cv4_spans(const double* rect, int W, int H, int32_t* rows, int cap) { return lsd_cv4_spans(rect, W, H, rows, cap); }
// stage outputs for the GPU stage-parity tests: blurred [h][w] u8, scaled [sh][sw] u8, modgrad / angles [sh][sw] double,
// order [(sw-1)(sh-1)] int32, region_id [sh][sw] int32.  Any pointer may be null.  Returns the number of segments.
int orc_lsd_stages(const uint8_t* img, int w, int h, int stride, int refine, uint8_t* blurred, uint8_t* scaled, double* modgrad, double* angles,
                   int32_t* order, int32_t* region_id, int32_t* sw_sh) {
    std::vector<LsdSegment> out;
    LsdStages st;
    lsd_detect_stages(Img8{img, w, h, stride}, refine, out, st);
    if (blurred) std::memcpy(blurred, st.blurred.data(), st.blurred.size());
    if (scaled) std::memcpy(scaled, st.scaled.data(), st.scaled.size());
    if (modgrad) std::memcpy(modgrad, st.modgrad.data(), st.modgrad.size() * 8);
    if (angles) std::memcpy(angles, st.angles.data(), st.angles.size() * 8);
    if (order) std::memcpy(order, st.order.data(), st.order.size() * 4);
    if (region_id) std::memcpy(region_id, st.region_id.data(), st.region_id.size() * 4);
    if (sw_sh) { sw_sh[0] = st.w; sw_sh[1] = st.h; }
    return (int)out.size();
}
// keylines: KeyLine[cap] (68 bytes each), lf: double[cap][3]; returns the number kept
int orc_extract_line_segments_enum(const uint8_t* img, int w, int h, int stride, int max_lines, int rect_enum, void* keylines, double* lf, int cap) {
    std::vector<KeyLine> kl; std::vector<double> f;
    extract_line_segments(Img8{img, w, h, stride}, max_lines, kl, f, rect_enum);
    const int n = std::min((int)kl.size(), cap);
    static_assert(sizeof(KeyLine) == 68, "KeyLine layout");
    if (n) { std::memcpy(keylines, kl.data(), (size_t)n * sizeof(KeyLine)); std::memcpy(lf, f.data(), (size_t)n * 3 * 8); }
    return n;
}
int orc_extract_line_segments(const uint8_t* img, int w, int h, int stride, int max_lines, void* keylines, double* lf, int cap) {
    return orc_extract_line_segments_enum(img, w, h, stride, max_lines, 0, keylines, lf, cap);
}
}

#include "linesearch.h"
extern "C" int orc_line_search_by_projection(int nf, const float* pt, const float* angle, const int32_t* octave, const uint8_t* desc, const uint8_t* has_obs,
                                             const float* scale_factors, int n_levels, int nm, const uint8_t* skip, const int32_t* level,
                                             const float* view_cos, const float* proj, const uint8_t* mdesc, const uint8_t* m_has_obs, float th,
                                             float nnratio, int32_t* assigned) {
    oracle::LineFrameView F; oracle::MapLinesView M;
    F.n = nf; F.pt = pt; F.angle = angle; F.octave = octave; F.desc = desc; F.has_obs = has_obs; F.scale_factors = scale_factors; F.n_levels = n_levels;
    M.n = nm; M.skip = skip; M.level = level; M.view_cos = view_cos; M.proj = proj; M.desc = mdesc; M.has_obs = m_has_obs;
    return oracle::line_search_by_projection(F, M, th, nnratio, assigned);
}

#include "bow.h"
extern "C" int orc_search_by_bow(int nkf, const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_has_mp, int kf_nodes, const int32_t* kf_node_id,
                                 const int32_t* kf_node_off, const int32_t* kf_node_feat, int nf, const uint8_t* f_desc, const float* f_angle, int f_nodes,
                                 const int32_t* f_node_id, const int32_t* f_node_off, const int32_t* f_node_feat, float nnratio, int check_ori, int32_t* match) {
    oracle::BowSide K, F;
    K.n = nkf; K.desc = kf_desc; K.angle = kf_angle; K.n_nodes = kf_nodes; K.node_id = kf_node_id; K.node_off = kf_node_off; K.node_feat = kf_node_feat;
    F.n = nf; F.desc = f_desc; F.angle = f_angle; F.n_nodes = f_nodes; F.node_id = f_node_id; F.node_off = f_node_off; F.node_feat = f_node_feat;
    return oracle::search_by_bow(K, kf_has_mp, F, nnratio, check_ori != 0, match);
}

#include "bow_transform.h"
// outputs sized by the caller: word_id / word_val [n], node_id [n], node_off [n + 1], node_feat [n], feat_word / feat_node [n]; counts[2] = {n_words, n_nodes}
extern "C" void orc_bow_transform(int n_nodes, int L, const uint8_t* vdesc, const int32_t* child_off, const int32_t* child_id, const int32_t* vword,
                                  const double* vweight, const uint8_t* features, int n, int levelsup, int32_t* word_id, double* word_val,
                                  int32_t* node_id, int32_t* node_off, int32_t* node_feat, int32_t* feat_word, int32_t* feat_node, int32_t* counts) {
    oracle::BowVocabulary V;
    V.n_nodes = n_nodes; V.L = L; V.desc = vdesc; V.child_off = child_off; V.child_id = child_id; V.word_id = vword; V.weight = vweight;
    oracle::BowResult R;
    oracle::bow_transform(V, features, n, levelsup, R);
    counts[0] = (int32_t)R.word_id.size(); counts[1] = (int32_t)R.node_id.size();
    std::memcpy(word_id, R.word_id.data(), R.word_id.size() * 4); std::memcpy(word_val, R.word_val.data(), R.word_val.size() * 8);
    std::memcpy(node_id, R.node_id.data(), R.node_id.size() * 4); std::memcpy(node_off, R.node_off.data(), R.node_off.size() * 4);
    std::memcpy(node_feat, R.node_feat.data(), R.node_feat.size() * 4);
    if (feat_word) std::memcpy(feat_word, R.feat_word.data(), (size_t)n * 4);
    if (feat_node) std::memcpy(feat_node, R.feat_node.data(), (size_t)n * 4);
}
extern "C" double orc_bow_score_l1(const int32_t* id1, const double* v1, int n1, const int32_t* id2, const double* v2, int n2) {
    return oracle::bow_score_l1(id1, v1, n1, id2, v2, n2);
}

// frame: Tcw[16], fx, fy, cx, cy, min_x, max_x, min_y, max_y, log_scale_factor (25 floats)
extern "C" void orc_lines_in_frustum(const float* frame, int n, const double* pos, const double* normal, const float* max_distance, const float* min_distance,
                                     float cos_limit, uint8_t* in_view, float* proj, int32_t* level, float* view_cos) {
    oracle::LineFrustumFrame F;
    for (int i = 0; i < 16; ++i) F.Tcw[i] = frame[i];
    F.fx = frame[16]; F.fy = frame[17]; F.cx = frame[18]; F.cy = frame[19]; F.min_x = frame[20]; F.max_x = frame[21]; F.min_y = frame[22]; F.max_y = frame[23];
    F.log_scale_factor = frame[24];
    oracle::lines_in_frustum(F, n, pos, normal, max_distance, min_distance, cos_limit, in_view, proj, level, view_cos);
}

#include "lbd.h"
extern "C" {
void orc_gaussian_blur_5x5_s1(const uint8_t* img, int w, int h, int stride, uint8_t* dst) { oracle::gaussian_blur_5x5_s1_u8(oracle::Img8{img, w, h, stride}, dst); }
void orc_sobel3_s16(const uint8_t* img, int w, int h, int16_t* dx, int16_t* dy) { oracle::sobel3_s16(img, w, h, dx, dy); }
// keylines: KeyLine[n] (68 bytes each); lbd72: float[n][72] or null; desc: uint8[n][32]
void orc_lbd_compute(const uint8_t* img, int w, int h, int stride, const void* keylines, int n, float* lbd72, uint8_t* desc) {
    oracle::lbd_compute(oracle::Img8{img, w, h, stride}, (const oracle::KeyLine*)keylines, n, lbd72, desc);
}
}

#include "detmath.h"
// oracle::det_sincos (detmath.h) on an array: what the product's lsd_detsincos.h must reproduce bit for bit
extern "C" void orc_det_sincos(const double* x, int n, double* s, double* c) {
    for (int i = 0; i < n; ++i) oracle::det_sincos(x[i], s[i], c[i]);
}
