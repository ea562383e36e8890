// DBoW2 vocabulary descent (TemplatedVocabulary::transform, Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1263) for all
// descriptors of a frame at once (it feeds SearchByBoW).  The tree lives in HBM as a CSR of children plus
// one 32-byte descriptor per node (ORBvoc: 10^6 leaves, ~35 MB); one warp walks one descriptor down the tree: the lanes
// take the children of the current node (k <= 32 per pass), one packed (distance << 8 | position) warp minimum per level
// picks the closest child, the first one on ties like `d < best_d` does.
#include <vector>
#include "fbe_internal.cuh"

struct fbe_vocabulary {
    int k = 0, L = 0, n_nodes = 0, device = 0;
    int* child_start = nullptr;   // [n_nodes + 2]
    int* child_items = nullptr;   // [n_nodes]
    uint8_t* desc = nullptr;      // [(n_nodes + 1) * 32]
    int* word_of = nullptr;       // [n_nodes + 1]
    double* weight = nullptr;     // [n_nodes + 1]
};

namespace fbe {

__global__ void __launch_bounds__(128) k_bow_transform(const int* __restrict__ child_start, const int* __restrict__ child_items,
                                                       const uint8_t* __restrict__ ndesc, const int* __restrict__ word_of,
                                                       const double* __restrict__ nweight, int nid_level,
                                                       const uint8_t* __restrict__ desc, int n, int* __restrict__ word_id,
                                                       int* __restrict__ node_id, double* __restrict__ weight) {
    const int f = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (f >= n) return;
    const uint4 q0 = *reinterpret_cast<const uint4*>(desc + (size_t)f * 32);
    const uint4 q1 = *reinterpret_cast<const uint4*>(desc + (size_t)f * 32 + 16);
    int cur = 0, level = 0, nid = 0;
    int beg = child_start[0], end = child_start[1];
    while (end > beg) {                                   // do { ... } while (!isLeaf()): the root of a non-empty tree has children
        ++level;
        unsigned best = 0xFFFFFFFFu;
        for (int base = beg; base < end; base += 32) {
            const int p = base + lane;
            unsigned key = 0xFFFFFFFFu;
            if (p < end) {
                const uint8_t* c = ndesc + (size_t)child_items[p] * 32;
                const uint4 c0 = *reinterpret_cast<const uint4*>(c), c1 = *reinterpret_cast<const uint4*>(c + 16);
                const int d = __popc(q0.x ^ c0.x) + __popc(q0.y ^ c0.y) + __popc(q0.z ^ c0.z) + __popc(q0.w ^ c0.w) +
                              __popc(q1.x ^ c1.x) + __popc(q1.y ^ c1.y) + __popc(q1.z ^ c1.z) + __popc(q1.w ^ c1.w);
                key = ((unsigned)d << 20) | (unsigned)(p - beg);
            }
            best = min(best, __reduce_min_sync(0xffffffffu, key));
        }
        cur = child_items[beg + (int)(best & 0xFFFFFu)];
        if (level == nid_level) nid = cur;
        beg = child_start[cur]; end = child_start[cur + 1];
    }
    if (lane == 0) { word_id[f] = word_of[cur]; node_id[f] = nid; weight[f] = nweight[cur]; }
}

}  // namespace fbe

using namespace fbe;

extern "C" {

int fbe_vocabulary_create(int32_t k, int32_t L, const int32_t* parent, const uint8_t* is_word, const uint8_t* desc,
                          const double* weight, int32_t n_nodes, int32_t device, fbe_vocabulary** out) {
    if (!out || n_nodes < 0 || (n_nodes > 0 && (!parent || !is_word || !desc || !weight))) return FBE_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { set_error("no CUDA device: this library has no CPU path"); return FBE_E_CUDA; }
    FBE_CUDA(cudaSetDevice(device));
    const int N = n_nodes + 1;                                        // + root
    std::vector<int> start(N + 1, 0), items(std::max(n_nodes, 1)), word(N, 0);
    std::vector<double> w(N, 0.0);
    std::vector<uint8_t> d((size_t)N * 32, 0);
    for (int i = 1; i <= n_nodes; ++i) {
        const int p = parent[i - 1];
        if (p < 0 || p >= i) { set_error("vocabulary: a node's parent must be created before it"); return FBE_E_INVALID; }
        start[p + 1]++;
    }
    for (int i = 0; i < N; ++i) start[i + 1] += start[i];
    std::vector<int> fill(start.begin(), start.end() - 1);
    int nwords = 0;
    for (int i = 1; i <= n_nodes; ++i) {
        items[fill[parent[i - 1]]++] = i;                             // children keep creation order (push_back, :1388)
        if (is_word[i - 1]) word[i] = nwords++;
        w[i] = weight[i - 1];
        for (int b = 0; b < 32; ++b) d[(size_t)i * 32 + b] = desc[(size_t)(i - 1) * 32 + b];
        if (start[parent[i - 1] + 1] - start[parent[i - 1]] >= (1 << 20)) { set_error("vocabulary: too many children"); return FBE_E_UNSUPPORTED; }
    }
    fbe_vocabulary* v = new fbe_vocabulary();
    v->k = k; v->L = L; v->n_nodes = n_nodes; v->device = device;
    bool ok = cudaMalloc(&v->child_start, (size_t)(N + 1) * 4) == cudaSuccess && cudaMalloc(&v->child_items, (size_t)std::max(n_nodes, 1) * 4) == cudaSuccess &&
              cudaMalloc(&v->desc, (size_t)N * 32) == cudaSuccess && cudaMalloc(&v->word_of, (size_t)N * 4) == cudaSuccess &&
              cudaMalloc(&v->weight, (size_t)N * 8) == cudaSuccess;
    ok = ok && cudaMemcpy(v->child_start, start.data(), (size_t)(N + 1) * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(v->child_items, items.data(), (size_t)std::max(n_nodes, 1) * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(v->desc, d.data(), (size_t)N * 32, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(v->word_of, word.data(), (size_t)N * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(v->weight, w.data(), (size_t)N * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) { fbe_vocabulary_destroy(v); set_error("vocabulary upload failed"); return FBE_E_CUDA; }
    *out = v;
    return FBE_OK;
}

int fbe_vocabulary_destroy(fbe_vocabulary* v) {
    if (!v) return FBE_E_INVALID;
    cudaFree(v->child_start); cudaFree(v->child_items); cudaFree(v->desc); cudaFree(v->word_of); cudaFree(v->weight);
    delete v;
    return FBE_OK;
}

int fbe_bow_transform(fbe_vocabulary* v, const uint8_t* desc, int32_t n, int32_t levelsup, int32_t* word_id, int32_t* node_id,
                      double* weight) {
    if (!v || n < 0 || (n > 0 && (!desc || !word_id || !node_id || !weight))) return FBE_E_INVALID;
    if (n == 0) return FBE_OK;
    if (v->n_nodes == 0) { set_error("empty vocabulary"); return FBE_E_INVALID; }     // transform(features,...) returns early on empty()
    FBE_CUDA(cudaSetDevice(v->device));
    uint8_t* d = nullptr;
    const size_t N = (size_t)n;
    FBE_CUDA(cudaMalloc(&d, N * 32 + N * 16 + 64));
    int* d_word = reinterpret_cast<int*>(d + N * 32);
    int* d_node = d_word + N;
    double* d_w = reinterpret_cast<double*>(d + N * 32 + ((N * 8 + 7) & ~(size_t)7));
    cudaError_t e = cudaMemcpy(d, desc, N * 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_bow_transform<<<(n + 3) / 4, 128>>>(v->child_start, v->child_items, v->desc, v->word_of, v->weight, v->L - levelsup, d, n,
                                              d_word, d_node, d_w);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(word_id, d_word, N * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(node_id, d_node, N * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(weight, d_w, N * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return FBE_E_CUDA; }
    return FBE_OK;
}

}  // extern "C"
