// ORBmatcher::DescriptorDistance (UPSTREAM src/ORBmatcher.cc) is a scalar host function on two cv::Mat rows; a single
// pair is not GPU work, so the shim keeps it inline (same integer as the SWAR original).  Batches of pairs, windowed
// candidate searches and brute-force kNN go to liborbx.so: orbx_distance_batch / orbx_match_windowed / orbx_knn2_*.
#ifndef ORBX_ORBMATCHER_DISTANCE_H
#define ORBX_ORBMATCHER_DISTANCE_H
#include <cstdint>
#include <cstring>

static inline int orbx_descriptor_distance(const unsigned char *a, const unsigned char *b) {
    int d = 0;
    for (int i = 0; i < 4; i++) {
        std::uint64_t x, y;
        std::memcpy(&x, a + 8 * i, 8); std::memcpy(&y, b + 8 * i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    return d;
}
#endif
