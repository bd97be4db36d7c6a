// Minimal stand-ins for the OpenCV types the shim touches, ONLY to compile-check shim/ORBextractor.cc in an image
// without OpenCV headers.  Never shipped: the reference image has the real <opencv2/opencv.hpp>.
#ifndef ORBX_CV_MIN_H
#define ORBX_CV_MIN_H
#include <cstddef>
#include <cstdlib>
#include <vector>
#define CV_8U 0
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_8UC4 24
namespace cv {
struct Point2f { float x, y; };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };
struct Mat {
    unsigned char *data = nullptr; int rows = 0, cols = 0; std::size_t step = 0; std::vector<unsigned char> store;
    int type() const { return CV_8UC1; }
    bool empty() const { return rows == 0 || cols == 0; }
    unsigned char *ptr(int r) { return data + (std::size_t)r * step; }
    void create(int r, int c, int) { rows = r; cols = c; step = (std::size_t)c; store.assign((std::size_t)r * c, 0); data = store.data(); }
    void release() { rows = cols = 0; data = nullptr; store.clear(); }
};
struct _InputArray { const Mat *m; _InputArray(const Mat &mm) : m(&mm) {} bool empty() const { return m->empty(); } Mat getMat() const { return *m; } };
struct _OutputArray { Mat *m; _OutputArray(Mat &mm) : m(&mm) {} void create(int r, int c, int t) const { m->create(r, c, t); } void release() const { m->release(); } Mat &getMat() const { return *m; } };
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;
}  // namespace cv
#endif
