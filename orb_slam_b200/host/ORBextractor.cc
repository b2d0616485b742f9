// ORBextractor.cc -- facade of ORB_SLAM::ORBextractor (reference src/ORBextractor.cc) over liborbfe.so.
//
// Replaces the reference translation unit wholesale: the constructor tables (:457-511), ComputePyramid
// (:781-822), ComputeKeyPoints (:522-707) and operator() (:718-779) all live behind the C-ABI
// (include/orbfe.h); this file only adapts cv:: types.
#include "ORBextractor.h"

#include <cstring>

#include <cassert>
#include <cstdio>
#include <cstdlib>

#include "orbfe.h"

namespace ORB_SLAM {

static int g_device = 0;
void ORBextractor::SetDevice(int device) { g_device = device; }

// Error policy, the same as ORBmatcher's (host/ORBmatcher.cc): the reference's extractor cannot fail, liborbfe's calls can
// (no device, a failed cudaMalloc, an unsupported geometry).  Every failure goes through one handler; the default logs and the
// call returns NO keypoints and an empty descriptor matrix (Tracking sees a frame without features: a lost frame, not a dead
// process); ORBFE_ABORT_ON_ERROR=1 or ORBextractor::SetErrorHandler() select another policy.  There is no CPU path.
static void default_error_handler(int code, const char *msg) {
    std::fprintf(stderr, "ORBextractor: liborbfe error %d: %s (there is no CPU path; returning no keypoints)\n", code, msg);
    const char *e = std::getenv("ORBFE_ABORT_ON_ERROR");
    if (e && *e && *e != '0') std::abort();
}
static ORBextractor::ErrorHandler g_error_handler = default_error_handler;
void ORBextractor::SetErrorHandler(ErrorHandler handler) { g_error_handler = handler ? handler : default_error_handler; }

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _scoreType, int _fastTh)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), scoreType(_scoreType), fastTh(_fastTh), mpImpl(NULL)
{
    const int rc = orbfe_extractor_create(_nfeatures, _scaleFactor, _nlevels, _scoreType, _fastTh, g_device, &mpImpl);
    if (rc != ORBFE_OK) {
        // the reference constructor cannot fail; there is deliberately no CPU path to fall back to: the object stays
        // without a handle and every operator() call reports through the handler again
        mpImpl = NULL;
        g_error_handler(rc, orbfe_last_error());
    }
}

ORBextractor::~ORBextractor() { if (mpImpl) orbfe_extractor_destroy(mpImpl); }

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors)
{
    if (_image.empty())  // reference :721-722: silently leave the outputs untouched
        return;
    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);  // :725
    (void)_mask;                      // never applied by the reference either (:601-603 build a cellMask nothing reads); empty at Frame.cc:60

    static_assert(sizeof(cv::KeyPoint) == sizeof(OrbfeKeyPoint), "cv::KeyPoint must be the 28-byte OpenCV 2.4 layout");
    const int cap = nfeatures > 0 ? nfeatures : 1;
    _keypoints.resize(cap);
    std::vector<unsigned char> desc((size_t)cap * 32);
    int n = 0;
    int rc = orbfe_extract(mpImpl, image.data, image.cols, image.rows, image.step,
                           reinterpret_cast<OrbfeKeyPoint*>(&_keypoints[0]), &desc[0], cap, &n);
    if (rc == ORBFE_ERR_CAPACITY) {  // sum of the level quotas exceeded nfeatures (:487 clips the last level at 0)
        _keypoints.resize(n);
        desc.resize((size_t)n * 32);
        rc = orbfe_extract(mpImpl, image.data, image.cols, image.rows, image.step,
                           reinterpret_cast<OrbfeKeyPoint*>(&_keypoints[0]), &desc[0], n, &n);
    }
    if (rc != ORBFE_OK) {
        g_error_handler(rc, orbfe_last_error());
        n = 0;
    }
    _keypoints.resize(n);
    if (n == 0) {
        _descriptors.release();  // :738-739
        return;
    }
    _descriptors.create(n, 32, CV_8U);  // :742
    cv::Mat out = _descriptors.getMat();
    for (int i = 0; i < n; i++) std::memcpy(out.ptr(i), &desc[(size_t)i * 32], 32);
}

}  // namespace ORB_SLAM
