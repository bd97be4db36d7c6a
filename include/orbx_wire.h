/* orbx_wire.h -- the SEND-SLAM wire messages either side of the ORB hot path, served by liborbx.so.
 *
 * The reference moves every camera frame from the Elixir app to the SLAM backend as a length-prefixed MessagePack map
 * (4-byte big-endian length + map; send_slam/lib/send_slam/slam_handler.ex:140-156,283-291) that carries the frame as a binary PPM:
 *     %{type: "frame", camera_id, encoding: "ppm", timestamp, width, height, channels, frame: <<bin>>}
 * and the backend parses it in ParseMessage (slam_backends/orb_slam_3/orbslam3_mono_networked.cc:302-337), decodes the PPM
 * (:546) and hands the Mat to TrackMonocular (:594).  The entry points below take the message bytes as they come off the socket
 * (the payload after the 4-byte length) so that imageData.assign / imdecode / cvtColor never run on the CPU, and define the
 * reverse message a consumer inside send_slam would send instead of a frame (keypoints + descriptors, ~60 KB instead of 2.7 MB).
 * Plain C, no allocation: pointers returned in the structs point INTO the payload the caller passed.
 */
#ifndef ORBX_WIRE_H
#define ORBX_WIRE_H

#include "orbx.h"

#ifdef __cplusplus
extern "C" {
#endif

/* What ParseMessage extracts from a non-calibration message (MessagePacket, orbslam3_mono_networked.cc:78-88). */
typedef struct orbx_wire_frame {
    const uint8_t *type;   /* value of "type" (not NUL-terminated), NULL if absent                          */
    size_t type_len;
    const uint8_t *image;  /* value of "frame" or "image" (must be a MessagePack bin), NULL if absent        */
    size_t image_bytes;
    double timestamp;      /* "timestamp": float32 / float64 / any integer, as msgpack-c's convert<double>   */
    int camera_id;         /* "camera_id": integer that fits an int                                          */
    int has_timestamp, has_camera_id;
} orbx_wire_frame;

/* Replaces ParseMessage (orbslam3_mono_networked.cc:302-337) for everything except the calibration section, which it skips
 * like any other unknown key: top level must be a map, keys must be strings, a later duplicate key overrides an earlier one,
 * "frame" / "image" that is not a bin is an error.  ORBX_OK, or ORBX_E_INVALID where ParseMessage throws (malformed or
 * truncated MessagePack, non-map root, wrong value types) or returns false (empty / missing type). */
int orbx_wire_parse_frame(const uint8_t *payload, size_t nbytes, orbx_wire_frame *out);

/* One iteration of the backend's receive loop for a "frame" message (orbslam3_mono_networked.cc:520-594) up to and including
 * ORBextractor::operator(): parse, the loop's own checks in its order (camera_id, image, timestamp present; image decodes),
 * then orbx_extract_pnm on the embedded PPM where it lies.  ORBX_E_EMPTY = the reference would log and skip this message
 * (reason in orbx_last_error); ORBX_E_INVALID = ParseMessage would throw, or the message is not a frame. */
int orbx_wire_process_frame(orbx_handle *h, const uint8_t *payload, size_t nbytes, int camera_rgb, int lap0, int lap1,
                            orbx_keypoint *kp_out, uint8_t *desc_out, int cap, int *n_out, int *mono_index_out, int *width_out,
                            int *height_out, double *timestamp_out, int *camera_id_out);

/* ---- the message of SURVEY.md §8f-3: features instead of a frame -----------------------------------------------------------
 * %{type: "features", camera_id, timestamp, width, height, mono_index, n, keypoints: <<n*28 B>>, descriptors: <<n*32 B>>}
 * keypoints = n cv::KeyPoint-compatible records (orbx_keypoint, little-endian floats), descriptors = n x 32 B.  Not part of the
 * reference's protocol: its C++ side would need one more branch next to `packet.type == "frame"` (:520) and an ORB-SLAM3 Frame
 * constructor that takes precomputed features.  framed != 0 prepends the 4-byte big-endian length the socket protocol uses. */
typedef struct orbx_wire_features {
    double timestamp;
    int camera_id, width, height, mono_index, n;
    const uint8_t *keypoints;         /* n x 28 bytes (orbx_keypoint records) at an ARBITRARY byte offset of the payload: never cast
                                         to orbx_keypoint *, copy them out (orbx_wire_copy_keypoints / memcpy) */
    const uint8_t *descriptors;       /* n x 32 bytes */
} orbx_wire_features;

size_t orbx_wire_features_bound(int n, int framed);   /* bytes orbx_wire_pack_features needs at most */
int orbx_wire_pack_features(double timestamp, int camera_id, int width, int height, int mono_index, const orbx_keypoint *kp,
                            const uint8_t *desc, int n, int framed, uint8_t *out, size_t out_cap, size_t *written);
int orbx_wire_parse_features(const uint8_t *payload, size_t nbytes, orbx_wire_features *out);
/* Copies the n keypoint records of a parsed features message into aligned caller memory (the std::vector<cv::KeyPoint> storage a
 * Frame would be built from).  Returns n, or ORBX_E_CAPACITY when cap < n. */
int orbx_wire_copy_keypoints(const orbx_wire_features *f, orbx_keypoint *dst, int cap);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_WIRE_H */
