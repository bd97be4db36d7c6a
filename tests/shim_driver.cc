// Test driver for seam b1: calls ORB_SLAM3::ORBextractor (shim/ORBextractor.cc on liborbx.so) the way UPSTREAM Frame::ExtractORB does,
// with the stand-in OpenCV types of shim/cv_min.h (this image has no OpenCV headers).  Built and called by tests/test_shim_gpu.py.
#include <cstring>
#include <vector>

#include "../shim/ORBextractor.h"

extern "C" {

void *shim_new(int nfeatures, float scale, int nlevels, int ini, int mn) { return new ORB_SLAM3::ORBextractor(nfeatures, scale, nlevels, ini, mn); }
void shim_delete(void *p) { delete static_cast<ORB_SLAM3::ORBextractor *>(p); }

// (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors, vLapping) of Frame::ExtractORB; returns monoIndex, *n_out = mvKeys.size()
int shim_extract(void *p, unsigned char *img, int w, int h, int stride, int lap0, int lap1, void *kps_out, unsigned char *desc_out, int cap,
                 int *n_out, int *desc_rows_out) {
    auto *ex = static_cast<ORB_SLAM3::ORBextractor *>(p);
    cv::Mat im; im.data = img; im.rows = h; im.cols = w; im.step = (std::size_t)stride;
    cv::Mat mask, desc;
    std::vector<cv::KeyPoint> keys;
    std::vector<int> lap = {lap0, lap1};
    const int mono = (*ex)(cv::_InputArray(im), cv::_InputArray(mask), keys, cv::_OutputArray(desc), lap);
    *n_out = (int)keys.size(); *desc_rows_out = desc.rows;
    if ((int)keys.size() <= cap) {
        std::memcpy(kps_out, keys.data(), keys.size() * sizeof(cv::KeyPoint));
        for (int i = 0; i < desc.rows; i++) std::memcpy(desc_out + (std::size_t)i * 32, desc.ptr(i), 32);
    }
    return mono;
}

int shim_getters(void *p, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2, float *scale_factor) {
    auto *ex = static_cast<ORB_SLAM3::ORBextractor *>(p);
    const int n = ex->GetLevels();
    *scale_factor = ex->GetScaleFactor();
    const auto a = ex->GetScaleFactors(), b = ex->GetInverseScaleFactors(), c = ex->GetScaleSigmaSquares(), d = ex->GetInverseScaleSigmaSquares();
    for (int i = 0; i < n; i++) { scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; }
    return n;
}

}  // extern "C"
