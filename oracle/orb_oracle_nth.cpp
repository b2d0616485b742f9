// orb_oracle_nth.cpp -- CPU ORACLE helper.  TEST INFRASTRUCTURE ONLY -- see orb_oracle.h.
//
// Literal cv::KeyPointsFilter::retainBest (OpenCV 2.4 features2d/keypoint.cpp semantics) followed by the
// reference's resize(n) (ORBextractor.cc:683-685, :697-701): std::nth_element at position n with a
// "response greater" comparator, then keep the first n.  Which of several equal-response keypoints
// survive is decided by libstdc++'s introselect; this mode exists to measure how often that differs
// from the canonical tie rule, not to define parity.
#include <algorithm>
#include "orb_oracle.h"

namespace {
struct ResponseGreater {
    bool operator()(const OrbOracleKeyPoint &a, const OrbOracleKeyPoint &b) const { return a.response > b.response; }
};
}  // namespace

extern "C" void orb_oracle_retain_best_nth(OrbOracleKeyPoint *kps, int *n_inout, int n_keep) {
    int n = *n_inout;
    if (n_keep < 0) n_keep = 0;
    if (n <= n_keep) return;
    if (n_keep == 0) { *n_inout = 0; return; }
    std::nth_element(kps, kps + n_keep, kps + n, ResponseGreater());
    // retainBest would now also keep every element equal to kps[n_keep-1].response (std::partition);
    // the reference immediately truncates to n_keep, so only the first n_keep survive.
    *n_inout = n_keep;
}
