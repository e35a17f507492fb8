// TEST INFRASTRUCTURE ONLY - see lbd.h.
#include "lbd.h"

#include <cmath>
#include <cstring>

#include "detmath.h"

namespace oracle {

void gaussian_blur_5x5_s1_u8(const Img8& s, uint8_t* dst) {
    static const unsigned k[5] = {14, 62, 104, 62, 14};
    const int w = s.w, h = s.h;
    std::vector<uint16_t> tmp((size_t)w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            unsigned acc = 0;
            for (int q = -2; q <= 2; ++q) acc += k[q + 2] * s.at(y, reflect101(x + q, w));
            tmp[(size_t)y * w + x] = (uint16_t)acc;
        }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int q = -2; q <= 2; ++q) acc += (uint32_t)k[q + 2] * tmp[(size_t)reflect101(y + q, h) * w + x];
            dst[(size_t)y * w + x] = (uint8_t)std::min<uint32_t>(255u, (acc + 32768u) >> 16);
        }
}

void sobel3_s16(const uint8_t* src, int w, int h, int16_t* dx, int16_t* dy) {
    auto at = [&](int y, int x) { return (int)src[(size_t)reflect101(y, h) * w + reflect101(x, w)]; };
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int a = at(y - 1, x - 1), b = at(y - 1, x), c = at(y - 1, x + 1), d = at(y, x - 1), f = at(y, x + 1), g = at(y + 1, x - 1), hh = at(y + 1, x),
                      i = at(y + 1, x + 1);
            dx[(size_t)y * w + x] = (int16_t)((c + 2 * f + i) - (a + 2 * d + g));
            dy[(size_t)y * w + x] = (int16_t)((g + 2 * hh + i) - (a + 2 * b + c));
        }
}

namespace {
const int kBands = 9, kWidth = 7, kHeight = kBands * kWidth;      // NUM_OF_BANDS, Params::widthOfBand_, heightOfLSP
// the 32 band pairs whose eight (mean, deviation) entries are compared
const int kComb[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7},
                          {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8}, {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};
}  // namespace

void lbd_gauss_tables(double* local21, double* global63) {
    {   // local weights F_l over three bands: centre (3 w - 1) / 2, sigma (2 w + 1) / 2 - both INTEGER divisions upstream
        const double u = (double)((kWidth * 3 - 1) / 2), sigma = (double)((kWidth * 2 + 1) / 2), inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < kWidth * 3; ++i) { const double dis = i - u; local21[i] = std::exp(dis * dis * inv); }
    }
    {   // global weights F_g over the whole support region
        const double u = (double)((kHeight - 1) / 2), sigma = u, inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < kHeight; ++i) { const double dis = i - u; global63[i] = std::exp(dis * dis * inv); }
    }
}

void lbd_compute(const Img8& img, const KeyLine* kl, int n, float* lbd72, uint8_t* desc) {
    const int w = img.w, h = img.h;
    std::vector<uint8_t> blur((size_t)w * h);
    gaussian_blur_5x5_s1_u8(img, blur.data());
    std::vector<int16_t> dxI((size_t)w * h), dyI((size_t)w * h);
    sobel3_s16(blur.data(), w, h, dxI.data(), dyI.data());
    double gl[kWidth * 3], gg[kHeight];
    lbd_gauss_tables(gl, gg);
    const short imageWidth = (short)(w - 1), imageHeight = (short)(h - 1), realWidth = (short)w;
    for (int li = 0; li < n; ++li) {
        const KeyLine& L = kl[li];
        float band[8][kBands];                 // pgdL, ngdL, pgdO, ngdO, pgdL2, ngdL2, pgdO2, ngdO2
        std::memset(band, 0, sizeof band);
        const short lengthOfLSP = (short)L.numOfPixels;
        const short halfWidth = (short)((lengthOfLSP - 1) / 2), halfHeight = (short)((kHeight - 1) / 2);
        const float midX = (float)(0.5 * (L.sPointInOctaveX + L.ePointInOctaveX)), midY = (float)(0.5 * (L.sPointInOctaveY + L.ePointInOctaveY));
        double sn, cs;
        det_sincos((double)L.angle, sn, cs);                       // cos / sin of the float direction (libm upstream; deterministic here, see detmath.h)
        const float dL0 = (float)cs, dL1 = (float)sn, dO0 = -dL1, dO1 = dL0;
        float sCorX0 = -dL0 * halfWidth + dL1 * halfHeight + midX, sCorY0 = -dL1 * halfWidth - dL0 * halfHeight + midY;
        for (short hID = 0; hID < kHeight; ++hID) {
            float sCorX = sCorX0, sCorY = sCorY0;
            float pL = 0, nL = 0, pO = 0, nO = 0;
            for (short wID = 0; wID < lengthOfLSP; ++wID) {
                short t = (short)std::round(sCorX);
                const short xCor = t < 0 ? 0 : (t > imageWidth ? imageWidth : t);
                t = (short)std::round(sCorY);
                const short yCor = t < 0 ? 0 : (t > imageHeight ? imageHeight : t);
                const short dx = dxI[(size_t)yCor * realWidth + xCor], dy = dyI[(size_t)yCor * realWidth + xCor];
                const float gDL = dx * dL0 + dy * dL1, gDO = dx * dO0 + dy * dO1;
                if (gDL > 0) pL += gDL; else nL -= gDL;
                if (gDO > 0) pO += gDO; else nO -= gDO;
                sCorX += dL0; sCorY += dL1;
            }
            sCorX0 -= dL1; sCorY0 += dL0;
            float coef = (float)gg[hID];
            pL = coef * pL; nL = coef * nL; pO = coef * pO; nO = coef * nO;
            const float row[8] = {pL, nL, pO, nO, pL * pL, nL * nL, pO * pO, nO * nO};
            auto add = [&](int b, float c) {
                for (int k = 0; k < 4; ++k) band[k][b] += c * row[k];
                for (int k = 4; k < 8; ++k) band[k][b] += c * c * row[k];
            };
            const int b = hID / kWidth, r = hID % kWidth;
            add(b, (float)gl[r + kWidth]);
            if (b - 1 >= 0) add(b - 1, (float)gl[r + 2 * kWidth]);
            if (b + 1 < kBands) add(b + 1, (float)gl[r]);
        }
        float des[kBands * 8];
        const float invN2 = (float)(1.0 / (kWidth * 2.0)), invN3 = (float)(1.0 / (kWidth * 3.0));
        for (int b = 0; b < kBands; ++b) {
            const float invN = (b == 0 || b == kBands - 1) ? invN2 : invN3;
            for (int k = 0; k < 4; ++k) {
                des[b * 8 + k] = band[k][b] * invN;
                const float t = band[4 + k][b] * invN - des[b * 8 + k] * des[b * 8 + k];
                des[b * 8 + 4 + k] = t > 0 ? std::sqrt(t) : 0.f;
            }
        }
        float tM = 0, tS = 0;
        for (int b = 0; b < kBands; ++b) {
            for (int k = 0; k < 4; ++k) tM += des[b * 8 + k] * des[b * 8 + k];
            for (int k = 4; k < 8; ++k) tS += des[b * 8 + k] * des[b * 8 + k];
        }
        tM = 1 / std::sqrt(tM); tS = 1 / std::sqrt(tS);
        for (int b = 0; b < kBands; ++b) {
            for (int k = 0; k < 4; ++k) des[b * 8 + k] *= tM;
            for (int k = 4; k < 8; ++k) des[b * 8 + k] *= tS;
        }
        for (int i = 0; i < kBands * 8; ++i) if (des[i] > 0.4) des[i] = (float)0.4;
        float t2 = 0;
        for (int i = 0; i < kBands * 8; ++i) t2 += des[i] * des[i];
        t2 = 1 / std::sqrt(t2);
        for (int i = 0; i < kBands * 8; ++i) des[i] = des[i] * t2;
        if (lbd72) std::memcpy(lbd72 + (size_t)li * 72, des, sizeof des);
        for (int c = 0; c < 32; ++c) {
            const float *f1 = des + 8 * kComb[c][0], *f2 = des + 8 * kComb[c][1];
            uint8_t v = 0;
            for (int i = 0; i < 8; ++i) if (f1[i] > f2[i]) v = (uint8_t)(v + (1u << (7 - i)));
            desc[(size_t)li * 32 + c] = v;
        }
    }
}

}  // namespace oracle
