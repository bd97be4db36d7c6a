// Vocabulary-tree descent of DBoW2: TemplatedVocabulary<ORB>::transform(feature, word_id, weight, nid, levelsup)
// as UPSTREAM Frame::ComputeBoW / KeyFrame::ComputeBoW call it through mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4)
// (ORB-SLAM3 src/Frame.cc; DBoW2 is an un-vendored third-party library of the reference build, slam_backends/orb_slam_3/CMakeLists.txt).
// Per descriptor: from the root, step to the child with the smallest Hamming distance (first child wins ties, `d < best_d`) until a
// leaf; report the leaf's word id and weight and the node reached `levelsup` levels above the leaves (the FeatureVector key).
// One warp per descriptor: lanes take the children of the current node (k = 10 in ORBvoc), a shuffle arg-min picks the next node;
// the 32-byte node descriptors are two 16-byte loads.  The walk is a chain of L dependent loads, so many descriptors in flight hide it.
#include <cuda_runtime.h>

#include <cstdint>
#include <new>
#include <string>
#include <vector>

#include "../../include/orbx.h"

struct orbx_vocab {
    int device = 0, n_nodes = 0, depth = 0;
    int32_t *d_child_start = nullptr, *d_child = nullptr, *d_word = nullptr;
    float *d_weight = nullptr;
    uint4 *d_desc = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
};

namespace {
thread_local std::string g_vocab_error;

__global__ void __launch_bounds__(256) k_bow_transform(const uint4 *__restrict__ feat, int n, const int32_t *__restrict__ child_start,
                                                       const int32_t *__restrict__ child, const uint4 *__restrict__ ndesc,
                                                       const int32_t *__restrict__ word, const float *__restrict__ weight, int nid_level,
                                                       int32_t *__restrict__ word_id, float *__restrict__ word_weight, int32_t *__restrict__ node_id) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const uint4 f0 = feat[2 * i], f1 = feat[2 * i + 1];
    int cur = 0, level = 0, nid = 0;
    while (true) {
        const int lo = child_start[cur], hi = child_start[cur + 1];
        if (hi <= lo) break;                                      // leaf
        level++;
        uint32_t best = 0xFFFFFFFFu;                              // (distance << 16 | position among the children): first minimum wins
        for (int c = lo + lane; c < hi; c += 32) {
            const int id = child[c];
            const uint4 a = ndesc[2 * id], b = ndesc[2 * id + 1];
            const int d = __popc(a.x ^ f0.x) + __popc(a.y ^ f0.y) + __popc(a.z ^ f0.z) + __popc(a.w ^ f0.w) + __popc(b.x ^ f1.x) +
                          __popc(b.y ^ f1.y) + __popc(b.z ^ f1.z) + __popc(b.w ^ f1.w);
            best = min(best, ((uint32_t)d << 16) | (uint32_t)(c - lo));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
        cur = child[lo + (int)(best & 0xFFFFu)];
        if (level == nid_level) nid = cur;
    }
    if (lane == 0) { word_id[i] = word[cur]; word_weight[i] = weight[cur]; node_id[i] = nid; }
}

#define V_TRY(v, expr)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) { (v)->err = std::string(#expr) + ": " + cudaGetErrorString(e__); return ORBX_E_CUDA; } \
    } while (0)
}  // namespace

extern "C" {

int orbx_vocab_create(int device, const int32_t *parent, const uint8_t *desc, const float *weight, int n_nodes, orbx_vocab **out) {
    if (!out) return ORBX_E_INVALID;
    *out = nullptr;
    if (!parent || !desc || !weight || n_nodes < 2 || parent[0] != -1) { g_vocab_error = "bad vocabulary (node 0 must be the root, parent -1)"; return ORBX_E_INVALID; }
    // children in node-id order (DBoW2 appends children in file order = id order); words are the leaves in id order
    std::vector<int32_t> start((size_t)n_nodes + 1, 0), child((size_t)n_nodes - 1), word((size_t)n_nodes, -1), depth((size_t)n_nodes, 0);
    for (int i = 1; i < n_nodes; i++) {
        if (parent[i] < 0 || parent[i] >= i || parent[i] >= n_nodes) { g_vocab_error = "parent ids must precede their children"; return ORBX_E_INVALID; }
        start[parent[i] + 1]++;
    }
    for (int i = 0; i < n_nodes; i++) { if (start[i + 1] > 65535) { g_vocab_error = "more than 65535 children"; return ORBX_E_INVALID; } start[i + 1] += start[i]; }
    std::vector<int32_t> fill(start.begin(), start.end() - 1);
    int maxd = 0;
    for (int i = 1; i < n_nodes; i++) { child[fill[parent[i]]++] = i; depth[i] = depth[parent[i]] + 1; if (depth[i] > maxd) maxd = depth[i]; }
    int nwords = 0;
    for (int i = 0; i < n_nodes; i++) if (start[i + 1] == start[i]) word[i] = nwords++;
    orbx_vocab *v = new (std::nothrow) orbx_vocab();
    if (!v) return ORBX_E_INVALID;
    v->device = device; v->n_nodes = n_nodes; v->depth = maxd;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void **)&v->d_child_start, sizeof(int32_t) * ((size_t)n_nodes + 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&v->d_child, sizeof(int32_t) * (size_t)n_nodes);
    if (e == cudaSuccess) e = cudaMalloc((void **)&v->d_word, sizeof(int32_t) * (size_t)n_nodes);
    if (e == cudaSuccess) e = cudaMalloc((void **)&v->d_weight, sizeof(float) * (size_t)n_nodes);
    if (e == cudaSuccess) e = cudaMalloc((void **)&v->d_desc, (size_t)32 * n_nodes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(v->d_child_start, start.data(), sizeof(int32_t) * ((size_t)n_nodes + 1), cudaMemcpyHostToDevice, v->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(v->d_child, child.data(), sizeof(int32_t) * ((size_t)n_nodes - 1), cudaMemcpyHostToDevice, v->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(v->d_word, word.data(), sizeof(int32_t) * (size_t)n_nodes, cudaMemcpyHostToDevice, v->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(v->d_weight, weight, sizeof(float) * (size_t)n_nodes, cudaMemcpyHostToDevice, v->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(v->d_desc, desc, (size_t)32 * n_nodes, cudaMemcpyHostToDevice, v->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream);
    if (e != cudaSuccess) {
        g_vocab_error = std::string("vocabulary upload: ") + cudaGetErrorString(e) + " (orbx has no CPU fallback)";
        orbx_vocab_destroy(v);
        return ORBX_E_CUDA;
    }
    *out = v;
    return ORBX_OK;
}

void orbx_vocab_destroy(orbx_vocab *v) {
    if (!v) return;
    cudaSetDevice(v->device);
    if (v->stream) { cudaStreamSynchronize(v->stream); cudaStreamDestroy(v->stream); }
    cudaFree(v->d_child_start); cudaFree(v->d_child); cudaFree(v->d_word); cudaFree(v->d_weight); cudaFree(v->d_desc);
    delete v;
}

const char *orbx_vocab_last_error(const orbx_vocab *v) { return v ? v->err.c_str() : g_vocab_error.c_str(); }
int orbx_vocab_depth(const orbx_vocab *v) { return v ? v->depth : ORBX_E_INVALID; }

int orbx_vocab_transform(orbx_vocab *v, const uint8_t *desc, int n, int levelsup, int32_t *word_id, float *word_weight, int32_t *node_id) {
    if (!v) return ORBX_E_INVALID;
    if (n < 0 || levelsup < 0 || (n > 0 && (!desc || !word_id || !word_weight || !node_id))) { v->err = "null argument"; return ORBX_E_INVALID; }
    if (n == 0) return ORBX_OK;
    V_TRY(v, cudaSetDevice(v->device));
    uint8_t *d_f = nullptr; int32_t *d_w = nullptr, *d_n = nullptr; float *d_wt = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_f, (size_t)32 * n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_w, sizeof(int32_t) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_n, sizeof(int32_t) * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_wt, sizeof(float) * (size_t)n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_f, desc, (size_t)32 * n, cudaMemcpyHostToDevice, v->stream);
    if (e == cudaSuccess) {
        // DBoW2: nid_level = m_L - levelsup; <= 0 means the root (node 0)
        k_bow_transform<<<(n + 7) / 8, 256, 0, v->stream>>>(reinterpret_cast<const uint4 *>(d_f), n, v->d_child_start, v->d_child, v->d_desc, v->d_word,
                                                            v->d_weight, v->depth - levelsup, d_w, d_wt, d_n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(word_id, d_w, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, v->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(word_weight, d_wt, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, v->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(node_id, d_n, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, v->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream);
    cudaFree(d_f); cudaFree(d_w); cudaFree(d_n); cudaFree(d_wt);
    if (e != cudaSuccess) { v->err = std::string("transform: ") + cudaGetErrorString(e); return ORBX_E_CUDA; }
    return ORBX_OK;
}

}  // extern "C"
