// T1: static combinatorial tables of the complete 3-skeleton on n vertices.
// Replaces ConstraintMatrices.create (rectifier.py:24-64): same lexicographic
// (itertools.combinations) simplex order, but every face id comes from the closed-form
// combinatorial-number-system rank instead of a linear search, and the dense 0/1 matrices are
// only materialised on request.
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace topo {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

int sm_count() {
    // per device: a process may drive more than one GPU
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = dev >= 0 && dev < 64 ? dev : 0;
    if (cached[slot] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[slot] = n > 0 ? n : 148;
    }
    return cached[slot];
}

int ensure_dynamic_smem(const void* kernel, size_t bytes) {
    // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: the opt-in is recorded per (device, kernel)
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_pair(dev, kernel);
    auto it = done.find(key);
    if (it != done.end() && it->second >= bytes) return TOPO_OK;
    TOPO_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    done[key] = bytes;
    return TOPO_OK;
}

namespace {

struct Binom {
    std::vector<std::vector<int64_t>> c;
    explicit Binom(int n) : c(n + 1, std::vector<int64_t>(6, 0)) {
        for (int i = 0; i <= n; ++i) {
            c[i][0] = 1;
            for (int j = 1; j <= 5; ++j) c[i][j] = (i == 0) ? 0 : c[i - 1][j - 1] + c[i - 1][j];
        }
    }
    int64_t operator()(int a, int b) const { return (a < 0 || b < 0 || b > 5 || a < b) ? 0 : c[a][b]; }
};

// lexicographic rank of the sorted k-subset v[0..k) of {0..n-1}
int64_t lex_rank(const Binom& C, int n, int k, const int* v) {
    int64_t r = C(n, k) - 1;
    for (int i = 0; i < k; ++i) r -= C(n - 1 - v[i], k - i);
    return r;
}

template <typename T>
int upload(topo_tables* t, const std::vector<T>& host, const T** dev_out) {
    *dev_out = nullptr;
    if (host.empty()) return TOPO_OK;
    void* p = nullptr;
    TOPO_CUDA(cudaMalloc(&p, host.size() * sizeof(T)));
    t->dev_blocks.push_back(p);
    TOPO_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dev_out = static_cast<const T*>(p);
    return TOPO_OK;
}

__global__ void face_matrix_kernel(const int* __restrict__ faces, int n_rows, int n_cols, int arity,
                                   float* __restrict__ out) {
    const int64_t total = static_cast<int64_t>(n_rows) * n_cols;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int row = static_cast<int>(i / n_cols), col = static_cast<int>(i % n_cols);
        float v = 0.f;
        for (int a = 0; a < arity; ++a) v = (faces[row * arity + a] == col) ? 1.f : v;
        out[i] = v;
    }
}

}  // namespace
}  // namespace topo

using namespace topo;

extern "C" int topo_version(void) { return 100; }
extern "C" const char* topo_last_error(void) { return topo::g_last_error.c_str(); }

extern "C" int topo_tables_create(int n, topo_tables** out) { return topo_tables_create_ex(n, 1, out); }

extern "C" int topo_tables_create_ex(int n, int upload_to_device, topo_tables** out) {
    TOPO_REQUIRE(out != nullptr, "out is null");
    TOPO_REQUIRE(n >= 1 && n <= 256, "n_vertices must be in [1, 256]");
    Binom C(n);
    TOPO_REQUIRE(C(n, 1) + C(n, 2) + C(n, 3) + C(n, 4) < (int64_t(1) << 30), "complex too large for int32 ids");

    topo_tables* t = new topo_tables();
    t->device = -1;
    if (upload_to_device) TOPO_CUDA(cudaGetDevice(&t->device));
    DeviceTables& d = t->d;
    d.n = n;
    d.off[0] = 0;
    for (int r = 0; r < 4; ++r) {
        d.cnt[r] = static_cast<int>(C(n, r + 1));
        d.off[r + 1] = d.off[r] + d.cnt[r];
    }
    for (int r = 0; r < 3; ++r) d.ncof[r] = std::max(0, n - 1 - r);

    // ---- simplices in lexicographic order, faces by rank arithmetic ----
    for (int r = 0; r < 4; ++r) {
        const int k = r + 1;
        t->h_verts[r].reserve(static_cast<size_t>(d.cnt[r]) * k);
        t->h_faces[r].reserve(r ? static_cast<size_t>(d.cnt[r]) * k : 0);
        if (d.cnt[r] == 0) continue;
        int v[4];
        for (int i = 0; i < k; ++i) v[i] = i;
        while (true) {
            for (int i = 0; i < k; ++i) t->h_verts[r].push_back(v[i]);
            if (r >= 1) {
                // omit the last vertex first: face ids come out ascending
                for (int omit = k - 1; omit >= 0; --omit) {
                    int f[3], m = 0;
                    for (int i = 0; i < k; ++i)
                        if (i != omit) f[m++] = v[i];
                    t->h_faces[r].push_back(static_cast<int>(lex_rank(C, n, k - 1, f)));
                }
            }
            int i = k - 1;
            while (i >= 0 && v[i] == n - k + i) --i;
            if (i < 0) break;
            ++v[i];
            for (int j = i + 1; j < k; ++j) v[j] = v[j - 1] + 1;
        }
    }

    // ---- cofaces: visiting the cofaces in ascending id keeps every list ascending ----
    for (int r = 0; r < 3; ++r) {
        const int w = d.ncof[r], k1 = r + 2;
        t->h_cofaces[r].assign(static_cast<size_t>(d.cnt[r]) * w, -1);
        std::vector<int> fill(d.cnt[r], 0);
        for (int s = 0; s < d.cnt[r + 1]; ++s)
            for (int a = 0; a < k1; ++a) {
                const int f = (r == 0) ? static_cast<int>(t->h_verts[1][s * 2 + a]) : t->h_faces[r + 1][s * k1 + a];
                t->h_cofaces[r][static_cast<size_t>(f) * w + fill[f]++] = s;
            }
    }
    // rank-1 "faces" are the edge's vertices
    t->h_faces[1].clear();
    for (int s = 0; s < d.cnt[1]; ++s) {
        t->h_faces[1].push_back(static_cast<int>(t->h_verts[1][s * 2]));
        t->h_faces[1].push_back(static_cast<int>(t->h_verts[1][s * 2 + 1]));
    }

    // ---- explicit adjacency neighbour lists (operator builder) ----
    d.adj_w[0] = std::max(0, n - 1);
    d.adj_w[1] = std::max(0, 2 * (n - 2));
    d.adj_w[2] = std::max(0, 3 * (n - 3));
    d.adj_w[3] = std::max(0, 4 * (n - 4));
    int64_t adj_entries = 0;
    for (int r = 0; r < 4; ++r) adj_entries += static_cast<int64_t>(d.cnt[r]) * d.adj_w[r];
    const bool build_adj = adj_entries <= (int64_t(1) << 26);
    std::vector<int> nbr[4], via[4];
    if (build_adj) {
        std::vector<std::pair<int, int>> row;
        for (int r = 0; r < 4; ++r) {
            const int w = d.adj_w[r];
            nbr[r].assign(static_cast<size_t>(d.cnt[r]) * w, -1);
            via[r].assign(static_cast<size_t>(d.cnt[r]) * w, -1);
            if (w == 0) continue;
            for (int s = 0; s < d.cnt[r]; ++s) {
                row.clear();
                if (r == 0) {
                    for (int j = 0; j < d.ncof[0]; ++j) {
                        const int e = t->h_cofaces[0][static_cast<size_t>(s) * d.ncof[0] + j];
                        const int a = t->h_faces[1][e * 2], b = t->h_faces[1][e * 2 + 1];
                        row.emplace_back(a == s ? b : a, e);
                    }
                } else if (r < 3) {
                    for (int j = 0; j < d.ncof[r]; ++j) {
                        const int c = t->h_cofaces[r][static_cast<size_t>(s) * d.ncof[r] + j];
                        for (int a = 0; a < r + 2; ++a) {
                            const int f = t->h_faces[r + 1][c * (r + 2) + a];
                            if (f != s) row.emplace_back(f, c);
                        }
                    }
                } else {
                    for (int a = 0; a < 4; ++a) {
                        const int f = t->h_faces[3][s * 4 + a];
                        for (int j = 0; j < d.ncof[2]; ++j) {
                            const int c = t->h_cofaces[2][static_cast<size_t>(f) * d.ncof[2] + j];
                            if (c != s) row.emplace_back(c, f);
                        }
                    }
                }
                std::sort(row.begin(), row.end());
                for (size_t j = 0; j < row.size(); ++j) {
                    nbr[r][static_cast<size_t>(s) * w + j] = row[j].first;
                    via[r][static_cast<size_t>(s) * w + j] = row[j].second;
                }
            }
        }
    } else {
        for (int r = 0; r < 4; ++r) d.adj_w[r] = -1;  // operator builder unavailable at this size
    }

    // ---- upload ----
    int rc = TOPO_OK;
    d.faces[0] = nullptr;
    if (!upload_to_device) {   // host-only object: table queries work, kernels refuse it
        for (int r = 0; r < 4; ++r) d.faces[r] = d.adj_nbr[r] = d.adj_via[r] = nullptr;
        for (int r = 0; r < 3; ++r) d.cofaces[r] = nullptr;
        *out = t;
        return TOPO_OK;
    }
    for (int r = 1; r < 4 && rc == TOPO_OK; ++r) rc = upload(t, t->h_faces[r], &d.faces[r]);
    for (int r = 0; r < 3 && rc == TOPO_OK; ++r) rc = upload(t, t->h_cofaces[r], &d.cofaces[r]);
    for (int r = 0; r < 4 && rc == TOPO_OK; ++r) {
        d.adj_nbr[r] = d.adj_via[r] = nullptr;
        if (build_adj) {
            rc = upload(t, nbr[r], &d.adj_nbr[r]);
            if (rc == TOPO_OK) rc = upload(t, via[r], &d.adj_via[r]);
        }
    }
    if (rc != TOPO_OK) {
        topo_tables_destroy(t);
        return rc;
    }
    *out = t;
    return TOPO_OK;
}

extern "C" void topo_tables_destroy(topo_tables* t) {
    if (!t) return;
    for (void* p : t->dev_blocks) cudaFree(p);
    delete t;
}

extern "C" int topo_tables_sizes(const topo_tables* t, int64_t counts[4], int64_t offsets[5]) {
    TOPO_REQUIRE(t && counts && offsets, "null argument");
    for (int r = 0; r < 4; ++r) counts[r] = t->d.cnt[r];
    for (int r = 0; r < 5; ++r) offsets[r] = t->d.off[r];
    return TOPO_OK;
}

extern "C" int topo_tables_simplex_vertices(const topo_tables* t, int rank, int64_t* host_out) {
    TOPO_REQUIRE(t && host_out, "null argument");
    TOPO_REQUIRE(rank >= 0 && rank <= 3, "rank out of range");
    std::copy(t->h_verts[rank].begin(), t->h_verts[rank].end(), host_out);
    return TOPO_OK;
}

extern "C" int topo_tables_faces(const topo_tables* t, int rank, int32_t* host_out) {
    TOPO_REQUIRE(t && host_out, "null argument");
    TOPO_REQUIRE(rank >= 1 && rank <= 3, "rank out of range");
    std::copy(t->h_faces[rank].begin(), t->h_faces[rank].end(), host_out);
    return TOPO_OK;
}

extern "C" int topo_tables_cofaces(const topo_tables* t, int rank, int32_t* host_out) {
    TOPO_REQUIRE(t && host_out, "null argument");
    TOPO_REQUIRE(rank >= 0 && rank <= 2, "rank out of range");
    std::copy(t->h_cofaces[rank].begin(), t->h_cofaces[rank].end(), host_out);
    return TOPO_OK;
}

extern "C" int topo_tables_face_matrix(const topo_tables* t, int rank, float* dev_out, topo_stream_t stream) {
    TOPO_REQUIRE(t && dev_out, "null argument");
    TOPO_REQUIRE_TABLES_DEVICE(t);
    TOPO_REQUIRE(rank >= 1 && rank <= 3, "rank out of range");
    const int rows = t->d.cnt[rank], cols = t->d.cnt[rank - 1];
    if (rows == 0 || cols == 0) return TOPO_OK;
    const int64_t total = static_cast<int64_t>(rows) * cols;
    const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 8));
    face_matrix_kernel<<<grid, 256, 0, as_stream(stream)>>>(t->d.faces[rank], rows, cols, rank + 1, dev_out);
    TOPO_LAUNCH_CHECK();
    return TOPO_OK;
}
