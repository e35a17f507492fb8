// TEST INFRASTRUCTURE ONLY.  C entry points onto the REFERENCE's own ORB extractor - src/ORBextractor.cc, compiled unmodified from
// /root/reference with the stand-ins of oracle/ref/shims/ (containers + the six OpenCV image primitives, which are the oracle's
// cv2-pinned restatements).  Built into oracle/_ref/liborb_ref.so by `make -C oracle ref`; pins oracle/orb.cc (tests/test_oracle_orb_ref.py).
//
// Allocator: DistributeOctTree sorts pair<count, ExtractorNode*> (src/ORBextractor.cc:684), so ties between equally populated nodes
// are broken by HEAP ADDRESS - the reference's result depends on the allocator.  mode 1 routes every allocation of this library
// through a monotonic arena (addresses grow in allocation order, nothing is reused), which turns "address order" into "creation
// order" - the convention the oracle and the CUDA path document; mode 0 keeps glibc malloc (whatever order tcache / fastbins give).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "ORBextractor.h"

namespace {
thread_local char* g_arena = nullptr;
thread_local size_t g_used = 0, g_cap = 0;
}
// bound to this library only (-Wl,-Bsymbolic): the process' other modules keep their operator new
void* operator new(size_t n) {
    if (g_arena) {
        const size_t a = (g_used + 15) & ~(size_t)15;
        if (a + n <= g_cap) { g_used = a + n; return g_arena + a; }
        throw std::bad_alloc();
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](size_t n) { return operator new(n); }
static inline bool in_arena(void* p) { return g_arena && (char*)p >= g_arena && (char*)p < g_arena + g_cap; }
void operator delete(void* p) noexcept { if (p && !in_arena(p)) std::free(p); }
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }

static int run(const uint8_t* gray, int w, int h, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, void* kps, uint8_t* desc, int cap) {
    Planar_SLAM::ORBextractor ext(nfeatures, scale_factor, nlevels, ini_th, min_th);
    cv::Mat img(h, w, CV_8UC1, (void*)gray);
    std::vector<cv::KeyPoint> k;
    cv::Mat d;
    ext(img, cv::Mat(), k, d);
    const int n = (int)k.size();
    static_assert(sizeof(cv::KeyPoint) == 28, "cv::KeyPoint layout");
    if (n <= cap && n > 0) { std::memcpy(kps, k.data(), (size_t)n * 28); for (int i = 0; i < n; ++i) std::memcpy(desc + 32 * (size_t)i, d.ptr(i), 32); }
    return n;
}

extern "C" {
// keypoints: 7 x 4 bytes each (x, y, size, angle, response, octave, class_id); desc: 32 bytes each.  Returns the number of key points.
// monotonic_alloc: 1 = arena allocator (see above), 0 = glibc malloc.
int ref_orb_extract(const uint8_t* gray, int w, int h, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int monotonic_alloc, void* kps,
                    uint8_t* desc, int cap) {
    char* arena = nullptr;
    if (monotonic_alloc) {
        g_cap = (size_t)1 << 30;                       // virtual; only touched pages are committed
        arena = (char*)std::malloc(g_cap);
        if (!arena) return -1;
        g_used = 0; g_arena = arena;
    }
    int n = -1;
    try { n = run(gray, w, h, nfeatures, scale_factor, nlevels, ini_th, min_th, kps, desc, cap); } catch (...) { n = -2; }
    g_arena = nullptr; g_cap = 0;
    std::free(arena);
    return n;
}
}
