// TEST INFRASTRUCTURE ONLY — C entry points (ctypes) onto the oracle ORB extractor.
#include "orb.h"
#include <cstring>
using namespace oracle;
extern "C" {
void* orc_orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
    OrbParams p; p.nfeatures = nfeatures; p.scale_factor = scale_factor; p.nlevels = nlevels;
    p.ini_th_fast = ini_th; p.min_th_fast = min_th;
    return new OrbExtractor(p);
}
void orc_orb_destroy(void* h) { delete (OrbExtractor*)h; }
// kps: cap x 28 B, desc: cap x 32 B; returns N (may exceed cap; only cap rows are written)
int orc_orb_extract(void* h, const uint8_t* gray, int w, int h_, int stride, KeyPoint* kps, uint8_t* desc, int cap) {
    OrbExtractor* e = (OrbExtractor*)h;
    std::vector<KeyPoint> k; std::vector<uint8_t> d;
    e->extract(Img8{gray, w, h_, stride}, k, d);
    int n = (int)k.size(), m = n < cap ? n : cap;
    if (m > 0) { std::memcpy(kps, k.data(), (size_t)m * sizeof(KeyPoint)); std::memcpy(desc, d.data(), (size_t)m * 32); }
    return n;
}
int orc_orb_level_size(void* h, int level, int* w, int* hh) {
    OrbExtractor* e = (OrbExtractor*)h;
    if (level < 0 || level >= (int)e->pyramid.size()) return -1;
    *w = e->pyramid[level].w; *hh = e->pyramid[level].h; return 0;
}
int orc_orb_level_pixels(void* h, int level, uint8_t* out) {
    OrbExtractor* e = (OrbExtractor*)h;
    if (level < 0 || level >= (int)e->pyramid.size()) return -1;
    std::memcpy(out, e->pyramid[level].px.data(), e->pyramid[level].px.size()); return 0;
}
// candidates of a level as int triples (x,y,score) relative to (16,16); returns count
int orc_orb_level_candidates(void* h, int level, int* xys, int cap) {
    OrbExtractor* e = (OrbExtractor*)h;
    if (level < 0 || level >= (int)e->candidates.size()) return -1;
    const auto& c = e->candidates[level];
    for (int i = 0; i < (int)c.size() && i < cap; ++i) { xys[3*i] = c[i].x; xys[3*i+1] = c[i].y; xys[3*i+2] = c[i].score; }
    return (int)c.size();
}
// keypoints of a level in level coordinates (before the final scale multiply), list order
int orc_orb_level_keypoints(void* h, int level, KeyPoint* out, int cap) {
    OrbExtractor* e = (OrbExtractor*)h;
    if (level < 0 || level >= (int)e->level_kps.size()) return -1;
    const auto& k = e->level_kps[level];
    for (int i = 0; i < (int)k.size() && i < cap; ++i) out[i] = k[i];
    return (int)k.size();
}
void orc_orb_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* quota, int* umax16) {
    OrbExtractor* e = (OrbExtractor*)h;
    for (int i = 0; i < e->prm.nlevels; ++i) {
        scale[i] = e->scale[i]; inv_scale[i] = e->inv_scale[i]; sigma2[i] = e->sigma2[i]; inv_sigma2[i] = e->inv_sigma2[i];
        quota[i] = e->features_per_level[i];
    }
    for (int i = 0; i < 16; ++i) umax16[i] = e->umax[i];
}
// steered BRIEF on an already-blurred image, for pinning against cv2.ORB.compute
void orc_orb_describe(void* h, const uint8_t* blurred, int w, int hh, float x, float y, float angle, uint8_t* out32) {
    OrbExtractor* e = (OrbExtractor*)h;
    OrbExtractor::Level B; B.w = w; B.h = hh; B.px.assign(blurred, blurred + (size_t)w * hh);
    KeyPoint k{}; k.x = x; k.y = y; k.angle = angle;
    e->describe(B, k, out32);
}
float orc_orb_ic_angle(void* h, const uint8_t* img, int w, int hh, int x, int y) {
    OrbExtractor* e = (OrbExtractor*)h;
    OrbExtractor::Level L; L.w = w; L.h = hh; L.px.assign(img, img + (size_t)w * hh);
    return e->ic_angle(L, x, y);
}
}
