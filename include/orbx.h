/* orbx -- B200-native ORB front end for SEND-SLAM: C ABI (drop-in boundary b2 of SURVEY.md §8b).
 *
 * One core library (liborbx.so, hand-written CUDA for sm_100a) serves the three seams of the reference:
 *   b1  C++ class seam   ORB_SLAM3::ORBextractor / ORBmatcher::DescriptorDistance, compiled by the reference
 *                        from src/ORBextractor.cc, src/ORBmatcher.cc (slam_backends/orb_slam_3/CMakeLists.txt:52-53,
 *                        headers :80-81) and reached from orbslam3_mono_networked.cc:594 (TrackMonocular);
 *                        shim/ORBextractor.{h,cc} forwards to this header.
 *   b3  BEAM seam        nif/orbx_nif.c (erl_nif) -- the Elixir side receives frames as
 *                        {:camera_frame,{:ok,opts}} (send_slam/lib/send_slam/camera_producer.ex:190-208) and today
 *                        ships them as PPM over TCP (send_slam/lib/send_slam/slam_handler.ex:59-88).
 * All entry points: plain pointers and sizes, caller-allocated outputs with explicit capacity, int status
 * (0 = ok, < 0 = ORBX_E_*), never throw / abort.  A handle is single-flight; distinct handles are independent
 * (own CUDA stream + workspace, bound to one device).  There is NO CPU fallback: without a usable CUDA device
 * orbx_create fails with ORBX_E_CUDA.
 */
#ifndef ORBX_H
#define ORBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBX_OK 0
#define ORBX_E_INVALID (-1)   /* bad argument (null pointer, size out of range, unsupported parameter)          */
#define ORBX_E_CUDA (-2)      /* CUDA runtime error or no device; text in orbx_last_error                       */
#define ORBX_E_CAPACITY (-3)  /* caller buffer / configured maximum too small                                   */
#define ORBX_E_EMPTY (-4)     /* empty image: the reference returns -1 from operator() (UPSTREAM assert/empty)  */
#define ORBX_E_OVERFLOW (-5)  /* internal candidate buffer overflow (cannot happen with the default sizing)     */

#define ORBX_MAX_LEVELS 16
#define ORBX_DESC_BYTES 32

typedef struct orbx_handle orbx_handle;   /* extractor (+ windowed matcher workspace)  */
typedef struct orbx_db orbx_db;           /* row shard of a descriptor database (kNN)  */

/* Replaces the five ORBextractor ctor arguments (UPSTREAM include/ORBextractor.h; values the reference passes:
 * orbslam3_mono_networked.cc:193-206) plus what a GPU workspace needs to be sized once. */
typedef struct orbx_config {
    int nfeatures;       /* ORBextractor.nFeatures  (reference: 1250; 5x for the initialisation extractor)   */
    float scale_factor;  /* ORBextractor.scaleFactor (1.2); supported range (1, 2)                            */
    int nlevels;         /* ORBextractor.nLevels (8); 1..ORBX_MAX_LEVELS                                      */
    int ini_th_fast;     /* ORBextractor.iniThFAST (20)                                                       */
    int min_th_fast;     /* ORBextractor.minThFAST (7)                                                        */
    int device;          /* CUDA device ordinal                                                               */
    int max_width;       /* largest frame the handle will see (<= 4095)                                       */
    int max_height;
    int max_batch;       /* frames per orbx_extract_batch call (1 for the per-frame seam)                     */
} orbx_config;

/* Binary-compatible with cv::KeyPoint (7 x 4 bytes) so the shim can memcpy into std::vector<cv::KeyPoint>. */
typedef struct orbx_keypoint {
    float x, y;        /* pt, image coordinates (level coordinates * mvScaleFactor[octave])   */
    float size;        /* (int)(31 * mvScaleFactor[octave])                                    */
    float angle;       /* IC_Angle, degrees [0,360)                                            */
    float response;    /* FAST score                                                           */
    int32_t octave;
    int32_t class_id;  /* -1                                                                   */
} orbx_keypoint;

/* ---- lifecycle ------------------------------------------------------------------------------------------- */
int orbx_create(const orbx_config *cfg, orbx_handle **out);
void orbx_destroy(orbx_handle *h);
/* Last error text of this handle (h == NULL: last error of a failed orbx_create / orbx_knn2_create_db on this thread). */
const char *orbx_last_error(const orbx_handle *h);
/* Upper bound of keypoints one frame can return (nfeatures + 3 per level + slack); size kp/desc buffers with it. */
int orbx_keypoint_capacity(const orbx_handle *h);

/* Replaces GetLevels/GetScaleFactor(s)/GetInverseScaleFactors/GetScaleSigmaSquares/GetInverseScaleSigmaSquares
 * (UPSTREAM include/ORBextractor.h getters, read by the Frame ctor).  Each array (may be NULL) gets nlevels entries.
 * Returns nlevels. */
int orbx_get_tables(const orbx_handle *h, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2,
                    int *features_per_level);
/* Level geometry for a w x h frame: widths/heights of the nlevels pyramid planes. Returns nlevels. */
int orbx_get_level_sizes(const orbx_handle *h, int width, int height, int *widths, int *heights);

/* ---- extraction ------------------------------------------------------------------------------------------ */
/* Replaces ORBextractor::operator()(image, mask, keypoints, descriptors, vLappingArea) for one CV_8UC1 frame in HOST
 * memory (mask ignored as upstream).  kp_out / desc_out: cap records / cap*32 bytes.  *n_out = keypoints written,
 * *mono_index_out = the value operator() returns.  Output order is the reference's (stereo slots filled from the
 * back for lap0 <= x <= lap1, mono slots from the front). */
int orbx_extract(orbx_handle *h, const uint8_t *gray, int width, int height, int stride, int lap0, int lap1,
                 orbx_keypoint *kp_out, uint8_t *desc_out, int cap, int *n_out, int *mono_index_out);

/* Same for `batch` equally sized HOST frames (frames[i] = first pixel of frame i).  Outputs are batch blocks of
 * cap records.  Host<->device copies are pipelined with the kernels on the handle's streams. */
int orbx_extract_batch(orbx_handle *h, const uint8_t *const *frames, int batch, int width, int height, int stride,
                       int lap0, int lap1, orbx_keypoint *kp_out, uint8_t *desc_out, int cap, int *n_out,
                       int *mono_index_out);

/* The same call in two halves for callers that stream batches (a camera group per handle): _submit queues the uploads,
 * kernels and downloads of one batch and returns without waiting; _collect waits for that batch and fills n_out /
 * mono_index_out (and, for pageable result buffers, kp_out / desc_out).  frames / kp_out / desc_out must stay valid and
 * untouched until _collect returns.  One batch may be in flight per handle (a second _submit returns ORBX_E_INVALID);
 * two handles used alternately from one thread overlap the upload of one batch with the kernels and download of the
 * other.  orbx_extract_batch == _submit + _collect.  There is no counterpart in the reference: its operator() is
 * synchronous (UPSTREAM Frame::ExtractORB), which is what orbx_extract keeps. */
int orbx_extract_batch_submit(orbx_handle *h, const uint8_t *const *frames, int batch, int width, int height, int stride,
                              int lap0, int lap1, orbx_keypoint *kp_out, uint8_t *desc_out, int cap);
int orbx_extract_batch_collect(orbx_handle *h, int *n_out, int *mono_index_out);

/* Frames as the reference puts them on the wire: SlamHandler encodes every frame with Evision.imencode(".ppm")
 * (send_slam/lib/send_slam/slam_handler.ex:275-277, message field `encoding: "ppm"` :147) and the backend decodes it with
 * cv::imdecode(packet.imageData, cv::IMREAD_UNCHANGED) (slam_backends/orb_slam_3/orbslam3_mono_networked.cc:546) before
 * TrackMonocular converts it to gray.  orbx_pnm_header follows OpenCV's PxM header reader (white space and '#' comments
 * before each number, exactly one byte consumed after each number); ORBX_E_EMPTY where imdecode returns an empty Mat (bad
 * magic / header, truncated payload: the reference logs and skips the frame, :547-551), ORBX_E_INVALID for variants that
 * decode but are not 8-bit binary (P1-P4, maxval > 255).  orbx_extract_pnm = imdecode + cvtColor + operator() with the
 * payload uploaded as it lies: camera_rgb is the Camera.RGB flag (`rgb: 1`, slam_handler.ex:222) that selects RGB2GRAY vs
 * BGR2GRAY on the decoded (BGR) Mat.  The frame size is returned through width_out / height_out (may be null). */
int orbx_pnm_header(const uint8_t *data, size_t nbytes, int *width, int *height, int *channels, size_t *payload_offset);
int orbx_extract_pnm(orbx_handle *h, const uint8_t *data, size_t nbytes, int camera_rgb, int lap0, int lap1,
                     orbx_keypoint *kp_out, uint8_t *desc_out, int cap, int *n_out, int *mono_index_out, int *width_out,
                     int *height_out);

/* Device-resident variant: frames already in HBM at d_frames + i*frame_stride_bytes (row pitch `stride`), results
 * left in HBM (d_kp_out: batch*cap records, d_desc_out: batch*cap*32 B, d_n_out / d_mono_out: batch ints).
 * Asynchronous on the handle's stream; call orbx_sync before reading results from another stream. */
int orbx_extract_batch_device(orbx_handle *h, const uint8_t *d_frames, size_t frame_stride_bytes, int batch, int width,
                              int height, int stride, int lap0, int lap1, orbx_keypoint *d_kp_out,
                              uint8_t *d_desc_out, int cap, int *d_n_out, int *d_mono_out);
int orbx_sync(orbx_handle *h);

/* Page-locked host memory for frames / results (cudaHostAlloc through the library, for hosts that do not link the CUDA runtime themselves:
 * the BEAM, a Go or Java process).  Page-locked buffers are DMA'd directly by orbx_extract_batch (no staging copy) and are what its
 * CUDA-graph replay keys on.  write_combined != 0 asks for write-combined memory: faster for the device to read over PCIe and fine for
 * upload-only frame buffers that the CPU fills sequentially, but slow for the CPU to read back -- never use it for result buffers. */
void *orbx_host_alloc(size_t bytes, int write_combined);
void orbx_host_free(void *p);

/* Pixel format of the frames handed to orbx_extract / orbx_extract_batch / orbx_extract_batch_device.
 * Replaces the cv::cvtColor(RGB2GRAY / BGR2GRAY / RGBA2GRAY / BGRA2GRAY) that UPSTREAM Tracking::GrabImageMonocular runs
 * on the CPU between cv::imdecode (orbslam3_mono_networked.cc:546) and the Frame constructor; which of RGB / BGR applies
 * is the caller's Camera.RGB flag (`rgb: 1`, send_slam/lib/send_slam/slam_handler.ex:222).  With a colour format `stride`
 * is the BYTE stride of a colour row and the conversion runs on the device straight into the level-0 plane.
 * gray_shift selects OpenCV's 8U fixed-point form: ORBX_GRAY_Q15 = (R*9798 + G*19235 + B*3735 + 2^14) >> 15 (bit-exact
 * against cv2 4.13, the pin of this repository), ORBX_GRAY_Q14 = (R*4899 + G*9617 + B*1868 + 2^13) >> 14 (older builds). */
#define ORBX_FMT_GRAY8 0
#define ORBX_FMT_RGB8 1
#define ORBX_FMT_BGR8 2
#define ORBX_FMT_RGBA8 3
#define ORBX_FMT_BGRA8 4
#define ORBX_GRAY_Q15 15
#define ORBX_GRAY_Q14 14
int orbx_set_input_format(orbx_handle *h, int format, int gray_shift);
/* The conversion alone (parity tests): src = colour image (format != GRAY8), dst = gray plane. */
int orbx_debug_gray(orbx_handle *h, const uint8_t *src, int width, int height, int stride, int format, int gray_shift,
                    uint8_t *dst, int dst_stride);
/* Issue all further work of this handle on the caller's CUDA stream (cudaStream_t; NULL = back to the handle's own
 * stream), e.g. a torch stream so that the caller's events bracket the kernels.  The legacy default stream has the
 * handle NULL as well: name it as cudaStreamLegacy ((cudaStream_t)0x1) to order this handle's work with it. */
int orbx_set_stream(orbx_handle *h, void *cuda_stream);
/* Number of kernel launches issued by this handle since creation (bench.py's gpu_launches). */
long long orbx_launch_count(const orbx_handle *h);

/* Per-stage device timing of the NEXT extract calls (CUDA events on the handle's stream between the kernels);
 * orbx_get_stage_times returns the last batch's milliseconds for {pyramid, blur, fast, quadtree, finalize, describe}. */
int orbx_set_profiling(orbx_handle *h, int enable);
int orbx_get_stage_times(orbx_handle *h, float *ms6);

/* ---- stage inspection (parity tests; valid after an extract call on the same handle) ------------------------- */
/* Copies pyramid level `level` of batch frame `frame` (blurred != 0: the Gaussian-blurred plane) to host. */
int orbx_debug_get_level(orbx_handle *h, int frame, int level, int blurred, uint8_t *out, int out_stride,
                         int *width_out, int *height_out);
/* FAST candidates of (frame, level) before the quadtree: triples (x, y, response) relative to (16,16), unordered. */
int orbx_debug_get_candidates(orbx_handle *h, int frame, int level, float *xyr_out, int cap, int *n_out);
/* Keypoints selected by the quadtree for (frame, level) in list order: (x, y, response, angle) in level coordinates. */
int orbx_debug_get_level_keypoints(orbx_handle *h, int frame, int level, float *xyra_out, int cap, int *n_out);
/* Stand-alone stages on caller data (HOST buffers), for per-kernel parity tests. */
int orbx_debug_resize(orbx_handle *h, const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw, int dh,
                      int dstride);
int orbx_debug_blur(orbx_handle *h, const uint8_t *src, int w, int ht, int sstride, uint8_t *dst, int dstride);
/* DistributeOctTree on caller candidates: keys = n x (x, y, response) relative to (minX,minY); out_idx = indices into
 * keys in final list order. */
int orbx_debug_octree(orbx_handle *h, const float *keys, int n, int minX, int maxX, int minY, int maxY, int N,
                      int *out_idx, int cap, int *n_out);
/* IC_Angle + steered BRIEF for caller keypoints (x, y integer-valued level coordinates) on a caller image / blurred
 * image; angle_in == NULL: compute angles from `img`, else use the given angles for the descriptors. */
int orbx_debug_describe(orbx_handle *h, const uint8_t *img, const uint8_t *blurred, int w, int ht, int stride,
                        const float *xy, int n, const float *angle_in, float *angle_out, uint8_t *desc_out);

/* ---- matching ----------------------------------------------------------------------------------------------- */
/* Replaces loops of ORBmatcher::DescriptorDistance (UPSTREAM src/ORBmatcher.cc): dist[i] = Hamming(a[i], b[i]),
 * HOST buffers of n x 32 bytes. */
int orbx_distance_batch(orbx_handle *h, const uint8_t *a, const uint8_t *b, int n, int32_t *dist_out);

/* Replaces the candidate loop of ORBmatcher::SearchByProjection / SearchForInitialization
 * (Frame::GetFeaturesInArea + arg-min DescriptorDistance): per query q (descriptor, window centre u,v, radius r,
 * minLevel, maxLevel) over the train keypoints of one frame (cv::KeyPoint records + descriptors, HOST buffers),
 * grid = 64 x 48 cells over bounds {minX, minY, maxX, maxY}.  Outputs per query: best / second-best train index
 * (-1 if none) and distance (256 if none); ties resolved in the reference's visiting order. */
int orbx_match_windowed(orbx_handle *h, const uint8_t *q_desc, const float *q_uvr, const int32_t *q_levels, int nq,
                        const orbx_keypoint *t_kp, const uint8_t *t_desc, int nt, const float *bounds4,
                        int32_t *best_idx, int32_t *best_dist, int32_t *second_idx, int32_t *second_dist);

/* Brute-force Hamming kNN, k = 2 (cv::BFMatcher(NORM_HAMMING).knnMatch): one orbx_db = one row shard resident in HBM.
 * row_offset = global index of the shard's first row (results carry global indices). */
int orbx_knn2_create_db(int device, const uint8_t *rows, long long nrows, long long row_offset, orbx_db **out);
/* Same, rows already in HBM on `device` (not copied; must outlive the db). */
int orbx_knn2_create_db_device(int device, const uint8_t *d_rows, long long nrows, long long row_offset, orbx_db **out);
void orbx_knn2_destroy_db(orbx_db *db);
const char *orbx_knn2_last_error(const orbx_db *db);
/* HOST queries (nq x 32 B) -> idx[nq*2] (global row, -1 if missing), dist[nq*2] (-1 if missing); ties -> lowest row. */
int orbx_knn2_query(orbx_db *db, const uint8_t *queries, int nq, int32_t *idx_out, int32_t *dist_out);
/* Device-resident: d_queries (nq x 32 B) -> d_packed_out[nq*2] = (dist << 32 | global row), 0xFFFF...F if missing.
 * Asynchronous on the db's stream; orbx_knn2_sync to wait. */
int orbx_knn2_query_device(orbx_db *db, const uint8_t *d_queries, int nq, unsigned long long *d_packed_out);
/* Top-2 merge of `nparts` partial results (layout [part][query][2], e.g. the NCCL all-gather of every rank's
 * orbx_knn2_query_device output) into d_packed_out[nq*2]; all pointers in HBM on the db's device. */
int orbx_knn2_merge_device(orbx_db *db, const unsigned long long *d_partials, int nparts, int nq,
                           unsigned long long *d_packed_out);
/* ---- multi-GPU brute-force kNN: row-sharded database, top-2 merge over NCCL (BASELINE config 4) -------------------------
 * One process per GPU.  Every rank holds one row shard as an orbx_db (row_offset = global index of its first row) and passes the SAME
 * queries; orbx_knn2_query_sharded* computes the local top-2 (orbx_knn2_query_device), exchanges the nq x 2 x 8-byte partials with
 * ONE ncclAllGather over NVLink (32 KB per rank at 2000 queries) on the db's stream and merges them on every rank (orbx_knn2_merge_device):
 * identical results on all ranks, ties to the lowest global row as cv::BFMatcher(NORM_HAMMING) does.  The reference has no
 * counterpart (its matcher sources, slam_backends/orb_slam_3/CMakeLists.txt:53, run on one CPU); this is the seam a backend that
 * relocalises against a large map database would link.
 * NCCL is bound at run time (dlopen of libnccl.so.2, or of the path in ORBX_NCCL_LIB): liborbx.so itself does not link it, and a
 * process that already has NCCL loaded (e.g. through torch) shares that copy.  orbx_comm wraps one communicator:
 *   orbx_comm_unique_id  ncclGetUniqueId on ONE rank; ship the 128 bytes to the others by any means (file, socket, MPI, torch)
 *   orbx_comm_create     ncclCommInitRank on `device` -- collective over all ranks
 *   orbx_comm_adopt      wrap a ncclComm_t the caller already owns (not destroyed with the orbx_comm)                              */
typedef struct orbx_comm orbx_comm;
#define ORBX_NCCL_ID_BYTES 128
int orbx_comm_unique_id(uint8_t *id_out /* ORBX_NCCL_ID_BYTES */);
int orbx_comm_create(int device, int rank, int nranks, const uint8_t *id /* ORBX_NCCL_ID_BYTES */, orbx_comm **out);
int orbx_comm_adopt(void *nccl_comm, int rank, int nranks, orbx_comm **out);
void orbx_comm_destroy(orbx_comm *c);
const char *orbx_comm_last_error(const orbx_comm *c);   /* c == NULL: last error of a failed create / unique_id on this thread */
/* d_queries (nq x 32 B, identical on every rank) -> d_packed_out[nq*2] = (dist << 32 | global row) over the WHOLE database, on every
 * rank.  Asynchronous on the db's stream (orbx_knn2_sync to wait).  Collective: every rank of the communicator must call it. */
int orbx_knn2_query_sharded_device(orbx_db *db, orbx_comm *comm, const uint8_t *d_queries, int nq, unsigned long long *d_packed_out);
/* Same with HOST queries / results (idx = global row, -1 if the database has fewer than two rows); waits for the result. */
int orbx_knn2_query_sharded(orbx_db *db, orbx_comm *comm, const uint8_t *queries, int nq, int32_t *idx_out, int32_t *dist_out);

/* ---- Frame post-extraction steps ---------------------------------------------------------------------------
 * What UPSTREAM ORB-SLAM3 src/Frame.cc runs between ORBextractor::operator() and the matchers, on the device:
 *   Frame::UndistortKeyPoints   = cv::undistortPoints(pts, pts, K, mDistCoef, cv::Mat(), K)  (5 iterations, double precision)
 *   Frame::ComputeImageBounds   = the same for the four image corners
 *   Frame::AssignFeaturesToGrid = Frame::PosInGrid on the FRAME_GRID_COLS x FRAME_GRID_ROWS = 64 x 48 grid
 * Camera = pinhole + radial-tangential, the values of the calibration message (orbslam3_mono_networked.cc:173-176:
 * Camera1.fx fy cx cy k1 k2 p1 p2 [k3]).  k1 == 0 means "no distortion" exactly as upstream tests mDistCoef.at<float>(0). */
typedef struct orbx_camera { float fx, fy, cx, cy, k1, k2, p1, p2, k3; } orbx_camera;
#define ORBX_GRID_COLS 64
#define ORBX_GRID_ROWS 48
#define ORBX_GRID_CELLS (ORBX_GRID_COLS * ORBX_GRID_ROWS)
/* cv::undistortPoints on n (x, y) pairs (host buffers). */
int orbx_undistort_points(orbx_handle *h, const float *xy, int n, const orbx_camera *cam, float *xy_out);
/* bounds4_out = mnMinX, mnMinY, mnMaxX, mnMaxY of Frame::ComputeImageBounds for a width x height image. */
int orbx_image_bounds(orbx_handle *h, const orbx_camera *cam, int width, int height, float *bounds4_out);
/* One frame, host buffers: kp_un_out[n] = keypoints with undistorted pt (mvKeysUn); cell_start_out[ORBX_GRID_CELLS + 1] and
 * cell_items_out[n] = mGrid as CSR, cell = posX * ORBX_GRID_ROWS + posY, items of a cell in push_back (index) order;
 * cell_start_out[ORBX_GRID_CELLS] = number of keypoints inside the grid. */
int orbx_frame_grid(orbx_handle *h, const orbx_keypoint *kp, int n, const orbx_camera *cam, const float *bounds4,
                    orbx_keypoint *kp_un_out, int32_t *cell_start_out, int32_t *cell_items_out);
/* The same for the device-resident results of orbx_extract_batch_device: d_kp [batch][cap], d_n [batch] ->
 * d_kp_un [batch][cap], d_cell_start [batch][ORBX_GRID_CELLS + 1], d_cell_items [batch][cap].  Asynchronous on the handle's stream. */
int orbx_frame_grid_batch_device(orbx_handle *h, const orbx_keypoint *d_kp, const int *d_n, int batch, int cap,
                                 const orbx_camera *cam, const float *bounds4, orbx_keypoint *d_kp_un, int32_t *d_cell_start,
                                 int32_t *d_cell_items);

/* orbx_match_windowed on data that never left HBM: queries and the train frame's mvKeysUn / descriptors / feature grid (the outputs of
 * orbx_extract_batch_device + orbx_frame_grid_batch_device for one frame) are device pointers; the candidate set of a query is read off
 * the grid cells of its window (Frame::GetFeaturesInArea) instead of testing every train keypoint.  Same results, same tie order.
 * Asynchronous on the handle's stream. */
int orbx_match_windowed_grid_device(orbx_handle *h, const uint8_t *d_q_desc, const float *d_q_uvr, const int32_t *d_q_levels, int nq,
                                    const orbx_keypoint *d_t_kp_un, const uint8_t *d_t_desc, const int32_t *d_cell_start,
                                    const int32_t *d_cell_items, const float *bounds4, int32_t *d_best_idx, int32_t *d_best_dist,
                                    int32_t *d_second_idx, int32_t *d_second_dist);
/* The same search for npairs (query frame, train frame) pairs of ONE batch in one launch -- e.g. frame i against frame i + 1 for a whole
 * batch of a camera sequence (UPSTREAM Tracking::TrackWithMotionModel -> ORBmatcher::SearchByProjection once per frame).  Every device array
 * is laid out [batch][cap] as orbx_extract_batch_device / orbx_frame_grid_batch_device leave it (d_cell_start: [batch][ORBX_GRID_CELLS + 1]);
 * d_q_* and the four outputs are indexed by the pair's query frame, the train arrays by its train frame; the number of queries of a
 * frame is read from d_n[frame] on the device (no host round trip between extraction and search).  pair_*_frame are HOST arrays. */
int orbx_match_windowed_grid_batch_device(orbx_handle *h, int npairs, const int32_t *pair_query_frame, const int32_t *pair_train_frame, int batch,
                                          int cap, const uint8_t *d_q_desc, const float *d_q_uvr, const int32_t *d_q_levels, const int32_t *d_n,
                                          const orbx_keypoint *d_t_kp_un, const uint8_t *d_t_desc, const int32_t *d_cell_start,
                                          const int32_t *d_cell_items, const float *bounds4, int32_t *d_best_idx, int32_t *d_best_dist,
                                          int32_t *d_second_idx, int32_t *d_second_dist);

/* ---- Vocabulary-tree descent ------------------------------------------------------------------------------------
 * DBoW2 TemplatedVocabulary<ORB>::transform as UPSTREAM Frame::ComputeBoW calls it (mpORBvocabulary->transform(desc, BowVec, FeatVec, 4)):
 * per descriptor the word (leaf) reached by stepping to the child of smallest Hamming distance (first child wins ties), its weight, and
 * the node `levelsup` levels above the leaf level (FeatureVector key; node 0 = root when levelsup >= depth).  The tree is handed over
 * as arrays, node 0 = root, parents before children (the order of an ORBvoc.txt file): parent[n], desc[n][32], weight[n].  Word ids are
 * the leaves in node order, as DBoW2::createWords assigns them.  BowVector accumulation / normalisation stays with the caller. */
typedef struct orbx_vocab orbx_vocab;
int orbx_vocab_create(int device, const int32_t *parent, const uint8_t *desc, const float *weight, int n_nodes, orbx_vocab **out);
void orbx_vocab_destroy(orbx_vocab *v);
const char *orbx_vocab_last_error(const orbx_vocab *v);
int orbx_vocab_depth(const orbx_vocab *v);
int orbx_vocab_transform(orbx_vocab *v, const uint8_t *desc, int n, int levelsup, int32_t *word_id, float *word_weight, int32_t *node_id);

/* Distance backend of a shard: ORBX_KNN_TENSOR (default) = descriptors expanded to {-1,+1} int8, q.d = 256 - 2H on
 * tcgen05.mma kind::i8 with the top-2 taken from TMEM; ORBX_KNN_POPC = XOR + POPC on the CUDA cores.  Identical results. */
#define ORBX_KNN_POPC 0
#define ORBX_KNN_TENSOR 1
#define ORBX_KNN_TENSOR_FP4 2   /* {-1,+1} as E2M1 nibbles, tcgen05.mma kind::mxf4 with unit block scales: half the operand bytes, twice the MMA rate */
int orbx_knn2_set_backend(orbx_db *db, int backend);
int orbx_knn2_sync(orbx_db *db);
int orbx_knn2_set_stream(orbx_db *db, void *cuda_stream);
long long orbx_knn2_launch_count(const orbx_db *db);

/* Geometry plan probe, needs NO GPU (used by the CPU-only tests): level sizes, FAST cells per level, per-level
 * quotas, quadtree roots / tabulated depth and the algorithmic bytes per frame of SURVEY.md §8(d).  Arrays may be NULL.
 * Returns nlevels or ORBX_E_INVALID. */
int orbx_plan_probe(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int width, int height,
                    int *widths, int *heights, int *ncells, int *quota, int *n_ini, int *depth0,
                    long long *algorithmic_bytes);

/* Library build identification: "orbx <version> sm_100a". */
const char *orbx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_H */
