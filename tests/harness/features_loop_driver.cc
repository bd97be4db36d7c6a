// Harness for the backend half of the `features` message (integration/orbslam3_mono_networked.features.patch): the body of the new
// receive-loop branch, compiled against liborbx.so WITHOUT ORB-SLAM3 / OpenCV.  It reads length-prefixed MessagePack messages
// (4-byte big-endian length + payload, the socket framing of slam_backends/orb_slam_3/orbslam3_mono_networked.cc:425-452) from a file,
// runs every `features` payload through orbx_wire_parse_features + orbx_wire_copy_keypoints into the two containers a Frame would be
// built from (std::vector<KeyPoint>, rows of 32 descriptor bytes) and prints one line per message, so that the Python test can
// compare against what it packed.  Skips what the branch would skip (parse failure, missing camera id).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "orbx_wire.h"

struct KeyPoint { float x, y, size, angle, response; int octave, class_id; };   // cv::KeyPoint's layout
static_assert(sizeof(KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint must be 7 x 4 bytes");

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int handled = 0, skipped = 0;
    for (;;) {
        uint8_t len4[4];
        if (std::fread(len4, 1, 4, f) != 4) break;
        const uint32_t n = ((uint32_t)len4[0] << 24) | ((uint32_t)len4[1] << 16) | ((uint32_t)len4[2] << 8) | (uint32_t)len4[3];
        if (n == 0) { std::puts("skip empty"); skipped++; continue; }
        std::vector<uint8_t> payload(n);
        if (std::fread(payload.data(), 1, n, f) != n) { std::puts("truncated"); break; }
        orbx_wire_features feat;
        if (orbx_wire_parse_features(payload.data(), payload.size(), &feat) != ORBX_OK) { std::puts("skip parse"); skipped++; continue; }
        if (!feat.camera_id) { std::puts("skip camera_id"); skipped++; continue; }
        std::vector<KeyPoint> keys((size_t)feat.n);
        if (orbx_wire_copy_keypoints(&feat, reinterpret_cast<orbx_keypoint *>(keys.data()), feat.n) != feat.n) { std::puts("skip copy"); skipped++; continue; }
        std::vector<uint8_t> descriptors((size_t)feat.n * 32);
        if (feat.n) std::memcpy(descriptors.data(), feat.descriptors, descriptors.size());
        // checksums over what tracking would receive
        double sx = 0, sy = 0, sa = 0; long so = 0; unsigned long sd = 0;
        for (const KeyPoint &k : keys) { sx += k.x; sy += k.y; sa += k.angle; so += k.octave; }
        for (uint8_t b : descriptors) sd = sd * 131u + b;
        std::printf("features camera=%d t=%.6f %dx%d mono=%d n=%d sx=%.3f sy=%.3f sa=%.3f so=%ld sd=%lu\n", feat.camera_id, feat.timestamp, feat.width,
                    feat.height, feat.mono_index, feat.n, sx, sy, sa, so, sd);
        handled++;
    }
    std::fclose(f);
    std::printf("done handled=%d skipped=%d\n", handled, skipped);
    return 0;
}
