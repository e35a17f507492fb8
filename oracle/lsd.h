// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Restatement of the line-segment front end the reference runs per frame:
//   LineSegment::ExtractLineSegment            src/LSDextractor.cpp:13-39 (detect, keep the 40 longest, line functions)
//   cv::line_descriptor::LSDDetector::detect   opencv_contrib 3.4.1 line_descriptor (NOT in /root/reference): one octave ->
//                                              the input image itself -> cv::createLineSegmentDetector(LSD_REFINE_ADV)
//                                              ->detect(); KeyLine fields from the segment end points
//   cv::LineSegmentDetector (LSD_REFINE_ADV)   OpenCV imgproc lsd.cpp (NOT in /root/reference): Gaussian 7x7 s=0.75 (u8
//                                              fixed point) + INTER_LINEAR_EXACT x0.8, level-line angles via fastAtan2,
//                                              1024-bin pseudo-ordering by std::sort, region growing, rectangle fit,
//                                              density refinement, NFA validation with rectangle improvement
// PINNED: the detector half is checked bit-for-bit (end points as float32, width, precision, log-NFA) against
// cv2.createLineSegmentDetector(cv2.LSD_REFINE_ADV / STD / NONE).detect of the in-container cv2 4.13 on synthetic frames
// (tests/test_oracle_lsd.py; REFINE_ADV with rect_enum = 3, see lsd_detect below).  OpenCV 3.4.1's own rect_nfa pixel
// enumeration (the reference's pinned version) is not obtainable here: parity unpinned for that detail.  The LBD descriptor half (BinaryDescriptor::compute) has no obtainable oracle here
//  and is not restated.
#pragma once
#include <cstdint>
#include <vector>

#include "cvprims.h"

namespace oracle {

struct LsdSegment {
    float x1, y1, x2, y2;      // cv::Vec4f as LineSegmentDetector::detect returns it (input-image pixels)
    double width, prec, nfa;   // nfa = -1 unless refine == ADV
};

// refine: 0 = LSD_REFINE_NONE, 1 = LSD_REFINE_STD, 2 = LSD_REFINE_ADV
// rect_enum (bit flags): bit 0 - which pixel enumeration the NFA validation (refine == 2) uses: 0 the published LSD rectangle
// iterator (the CUDA path of this round), 1 cv2 4.13's; bit 1 - rectangle axes from the host libm instead of detmath.h.
// rect_enum = 3 reproduces cv2 4.13 LSD_REFINE_ADV in every field; see lsd.cc rect_nfa
void lsd_detect(const Img8& img, int refine, std::vector<LsdSegment>& out, int rect_enum = 0);

// row spans {y, xa, xb} of the cv2 4.13 rect_nfa enumeration for rect = {x1, y1, x2, y2, width, dx, dy} in a W x H image (for the host
// check of the CUDA path's span header); returns the number of non-empty rows
int lsd_cv4_spans(const double* rect, int W, int H, int32_t* rows, int cap);

// intermediate products (debug / GPU stage parity)
struct LsdStages {
    int w = 0, h = 0;                       // scaled image size
    std::vector<uint8_t> blurred;           // input size
    std::vector<uint8_t> scaled;            // w x h
    std::vector<double> modgrad, angles;    // w x h (last row / column: angle NOTDEF, modgrad 0)
    std::vector<int32_t> order;             // pixel indices y * w + x of the (w-1)(h-1) gradient pixels in seed order
    std::vector<int32_t> region_id;         // w x h: index of the accepted segment that owns the pixel at the end, -1 otherwise
};
void lsd_detect_stages(const Img8& img, int refine, std::vector<LsdSegment>& out, LsdStages& st, int rect_enum = 0);

// cv::line_descriptor::KeyLine (68 bytes, 17 four-byte fields) as LSDDetector::detectImpl fills it for octave 0
struct KeyLine {
    float angle; int32_t class_id; int32_t octave; float pt_x, pt_y; float response; float size;
    float startPointX, startPointY, endPointX, endPointY, sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
    float lineLength; int32_t numOfPixels;
};
// LineSegment::ExtractLineSegment without the LBD descriptors: detect, sort by response (descending, std::sort like the
// reference), keep max_lines, renumber class_id, line functions l = sp x ep / |sp x ep|.
void extract_line_segments(const Img8& img, int max_lines, std::vector<KeyLine>& kl, std::vector<double>& line_functions /* [n][3] */,
                           int rect_enum = 0);

}  // namespace oracle
