// ORB_SLAM3::ORBextractor on top of liborbx.so.  Replaces src/ORBextractor.cc in the reference's library build
// (slam_backends/orb_slam_3/CMakeLists.txt:52); see INTEGRATION.md for the two-line CMake change.
#include "ORBextractor.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../include/orbx.h"

namespace ORB_SLAM3 {

static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint must be 7 x 4 bytes");

static std::atomic<int> g_device{-1};   // -1: not chosen by the program -> ORBX_DEVICE, else 0
void ORBextractor::SetDevice(int cuda_device) { g_device.store(cuda_device); }
int ORBextractor::GetDevice() {
    const int d = g_device.load();
    if (d >= 0) return d;
    const char *e = std::getenv("ORBX_DEVICE");
    return e && *e ? std::atoi(e) : 0;
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
    orbx_config cfg;
    cfg.nfeatures = _nfeatures; cfg.scale_factor = _scaleFactor; cfg.nlevels = _nlevels;
    cfg.ini_th_fast = _iniThFAST; cfg.min_th_fast = _minThFAST;
    cfg.device = GetDevice(); cfg.max_width = 4095; cfg.max_height = 4095; cfg.max_batch = 1;   // workspace is sized on first use
    if (orbx_create(&cfg, &mHandle) != ORBX_OK) {
        // the reference never throws here; keep the process alive and fail every call loudly instead
        std::fprintf(stderr, "ORBextractor(orbx): %s\n", orbx_last_error(nullptr));
        mHandle = nullptr;
        return;
    }
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels); mnFeaturesPerLevel.resize(nlevels);
    orbx_get_tables(mHandle, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                    mnFeaturesPerLevel.data());
    mCap = orbx_keypoint_capacity(mHandle);
    mDesc.resize((size_t)mCap * ORBX_DESC_BYTES);
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() { orbx_destroy(mHandle); }

bool ORBextractor::SetInputFormat(int orbx_fmt, int gray_shift) {
    if (!mHandle || orbx_set_input_format(mHandle, orbx_fmt, gray_shift) != ORBX_OK) return false;
    mInputFormat = orbx_fmt;
    return true;
}

int ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint> &_keypoints,
                             cv::OutputArray _descriptors, std::vector<int> &vLappingArea) {
    if (_image.empty() || !mHandle) return -1;
    cv::Mat image = _image.getMat();
    // upstream: assert(image.type() == CV_8UC1).  Extension: after SetInputFormat(ORBX_FMT_RGB8 ...) the colour
    // Mat that Tracking::GrabImageMonocular would have run through cv::cvtColor is accepted as it is.
    const int want = mInputFormat == ORBX_FMT_GRAY8 ? CV_8UC1 : mInputFormat >= ORBX_FMT_RGBA8 ? CV_8UC4 : CV_8UC3;
    if (image.type() != want) return -1;
    const int lap0 = vLappingArea.size() > 0 ? vLappingArea[0] : 0, lap1 = vLappingArea.size() > 1 ? vLappingArea[1] : 0;
    _keypoints.resize((size_t)mCap);
    int n = 0, mono = -1;
    const int rc = orbx_extract(mHandle, image.data, image.cols, image.rows, (int)image.step, lap0, lap1,
                                reinterpret_cast<orbx_keypoint *>(_keypoints.data()), mDesc.data(), mCap, &n, &mono);
    if (rc != ORBX_OK) {
        std::fprintf(stderr, "ORBextractor(orbx): %s\n", orbx_last_error(mHandle));
        _keypoints.clear(); _descriptors.release();
        return -1;
    }
    _keypoints.resize((size_t)n);
    if (n == 0) { _descriptors.release(); return mono; }
    _descriptors.create(n, 32, CV_8U);
    cv::Mat descriptors = _descriptors.getMat();
    for (int i = 0; i < n; i++) std::memcpy(descriptors.ptr(i), mDesc.data() + (size_t)i * 32, 32);
    return mono;
}

}  // namespace ORB_SLAM3
