// Drop-in replacement for ORB-SLAM3's include/ORBextractor.h (the header the reference installs per
// slam_backends/orb_slam_3/CMakeLists.txt:80 and whose class Frame/Tracking use behind
// orbslam3_mono_networked.cc:594).  Same class name, constructor, operator(), getters and public members; the body
// forwards to the C ABI of liborbx.so (include/orbx.h).  Built against real OpenCV headers in the reference image;
// here it is compile-checked against shim/cv_min.h (-DORBX_SHIM_CV_MIN).
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <list>
#include <vector>
#ifdef ORBX_SHIM_CV_MIN
#include "cv_min.h"
#else
#include <opencv2/opencv.hpp>
#endif

struct orbx_handle;

namespace ORB_SLAM3 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor &) = delete;
    ORBextractor &operator=(const ORBextractor &) = delete;

    // Compute the ORB features and descriptors on an image.  Mask is ignored (as upstream).  Returns monoIndex,
    // -1 for an empty image.
    int operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint> &_keypoints,
                   cv::OutputArray _descriptors, std::vector<int> &vLappingArea);

    // not in upstream: colour frames converted on the device (cv::cvtColor *2GRAY of Tracking::GrabImageMonocular)
    bool SetInputFormat(int orbx_fmt, int gray_shift = 15 /* ORBX_GRAY_Q15 */);

    // not in upstream: the CUDA device the extractors constructed AFTER this call live on (System constructs them inside Tracking, so
    // a backend that wants GPU 3 calls ORB_SLAM3::ORBextractor::SetDevice(3) before `new ORB_SLAM3::System(...)`,
    // orbslam3_mono_networked.cc:511).  Without a call the environment variable ORBX_DEVICE decides, else device 0.
    static void SetDevice(int cuda_device);
    static int GetDevice();

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // INTENTIONALLY EMPTY Mats (nlevels of them): upstream fills this with the bordered pyramid planes, and only the stereo matcher
    // (Frame::ComputeStereoMatches) reads it.  The mono path of SEND-SLAM (System::MONOCULAR, orbslam3_mono_networked.cc:511) never
    // does, and the planes live in HBM here; a stereo port would download them with orbx_debug_get_level.
    std::vector<cv::Mat> mvImagePyramid;

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;

private:
    int mInputFormat = 0;                // ORBX_FMT_GRAY8
    orbx_handle *mHandle = nullptr;      // one CUDA stream + workspace; single-flight like the reference's call pattern
    std::vector<unsigned char> mDesc;    // staging for descriptors (cap x 32)
    int mCap = 0;
};

}  // namespace ORB_SLAM3

#endif
