// Periodic radius graph on the GPU, bit-exact against the reference's radius_graph_pbc
// (/root/reference/dataset/utils.py:57-237) in edge set AND order.
//
// Design (B200-first, not a port of the reference's O(n^2 C) tensor expression):
//   * one warp per destination atom i1; lanes take 32 consecutive sources i2, so the output order
//     (dst, src, cell) falls out of the traversal itself -- no sort, no atomics;
//   * per (i1,i2) pair only the lattice images that can possibly be within the radius are visited:
//     the fractional-coordinate separation bounds the image index along each axis to
//     [df_k - r|b_k|, df_k + r|b_k|] (b_k = reciprocal vectors), clipped to the reference's +-rep_k
//     search range. At ADP density that is ~1 candidate per pair instead of 27;
//   * every candidate is then tested with the reference's exact fp32 operation sequence
//     (separately rounded products, no FMA contraction; the bmm summation order is the one ATen
//     uses for that number of cells), so the accepted set is identical bit for bit;
//   * count pass -> scan -> fill pass; the caller allocates the edge arrays in between.
#include "common.cuh"

namespace cartnet {

struct CellGeom {
    float c[9];      // cell rows (lattice vectors), fp32 as given
    double inv[9];   // inverse: frac = pos @ inv
    double rho[3];   // conservative half-range of the image index along each axis
    int rep[3];
    int mode;        // 0: (t0+t1)+t2, 1: (t0+t2)+t1  (see oracle.offset_sum_mode)
};

__device__ __forceinline__ void load_geom(CellGeom& g, const float* __restrict__ cell, const int32_t* reps,
                                          int reps_stride, int b, float radius) {
#pragma unroll
    for (int i = 0; i < 9; ++i) g.c[i] = cell[b * 9 + i];
    const int32_t* r = reps + (int64_t)b * reps_stride;
    g.rep[0] = r[0]; g.rep[1] = r[1]; g.rep[2] = r[2];
    int C = (2 * g.rep[0] + 1) * (2 * g.rep[1] + 1) * (2 * g.rep[2] + 1);
    g.mode = C < 45 ? 0 : 1;
    double a[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) a[i] = (double)g.c[i];
    double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    double id = 1.0 / det;
    // inverse = adj / det ; adj[i][j] = cofactor[j][i]
    g.inv[0] = c00 * id;
    g.inv[1] = (a[2] * a[7] - a[1] * a[8]) * id;
    g.inv[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    g.inv[3] = c01 * id;
    g.inv[4] = (a[0] * a[8] - a[2] * a[6]) * id;
    g.inv[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    g.inv[6] = c02 * id;
    g.inv[7] = (a[1] * a[6] - a[0] * a[7]) * id;
    g.inv[8] = (a[0] * a[4] - a[1] * a[3]) * id;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double bn = sqrt(g.inv[k] * g.inv[k] + g.inv[3 + k] * g.inv[3 + k] + g.inv[6 + k] * g.inv[6 + k]);
        g.rho[k] = (double)radius * bn * 1.001 + 1e-3;   // margin >> fp32 rounding of the exact test
    }
}

// The reference's fp32 test for one (pair, image). Returns d^2 and the separation vector.
__device__ __forceinline__ float exact_d2(const CellGeom& g, const float p1[3], const float p2[3], int u1, int u2,
                                          int u3, float delta[3]) {
    float fu1 = (float)u1, fu2 = (float)u2, fu3 = (float)u3;
    float sq[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float t0 = __fmul_rn(g.c[0 + d], fu1);
        float t1 = __fmul_rn(g.c[3 + d], fu2);
        float t2 = __fmul_rn(g.c[6 + d], fu3);
        float off = g.mode == 0 ? __fadd_rn(__fadd_rn(t0, t1), t2) : __fadd_rn(__fadd_rn(t0, t2), t1);
        float q = __fadd_rn(p2[d], off);            // utils.py:193
        delta[d] = __fsub_rn(p1[d], q);             // utils.py:196
        sq[d] = __fmul_rn(delta[d], delta[d]);
    }
    return __fadd_rn(__fadd_rn(sq[0], sq[1]), sq[2]);  // utils.py:197
}

struct NlistOut {
    int64_t* edge_index; int64_t num_edges;
    float *unit_cell, *dist, *direction, *cart_dist, *cart_dir;
    int32_t *src32, *dst32;
};

// one accepted (pair, image): every output field of edge w (utils.py:206-213,235-237; figshare_dataset.py:67-68)
__device__ __forceinline__ void write_edge(const NlistOut& o, int64_t w, int i1, int i2, int u1, int u2, int u3, float d2,
                                           const float delta[3]) {
    o.edge_index[w] = (int64_t)i2;                 // row 0: source j   (utils.py:235)
    o.edge_index[o.num_edges + w] = (int64_t)i1;   // row 1: destination i
    o.unit_cell[3 * w + 0] = (float)u1;
    o.unit_cell[3 * w + 1] = (float)u2;
    o.unit_cell[3 * w + 2] = (float)u3;
    const float dd = __fsqrt_rn(d2);
    o.dist[w] = dd;
    o.direction[3 * w + 0] = delta[0];
    o.direction[3 * w + 1] = delta[1];
    o.direction[3 * w + 2] = delta[2];
    if (o.cart_dist) o.cart_dist[w] = dd;            // figshare_dataset.py:67
    if (o.cart_dir) {                                // figshare_dataset.py:68
        const float nrm = fmaxf(dd, 1e-12f);
        o.cart_dir[3 * w + 0] = __fdiv_rn(delta[0], nrm);
        o.cart_dir[3 * w + 1] = __fdiv_rn(delta[1], nrm);
        o.cart_dir[3 * w + 2] = __fdiv_rn(delta[2], nrm);
    }
    if (o.src32) o.src32[w] = i2;
    if (o.dst32) o.dst32[w] = i1;
}

// All-pairs scan of one destination row by one warp (small crystals): returns the row's edge count.
template <bool FILL>
__device__ __forceinline__ int scan_all_pairs(const CellGeom& g, const float* __restrict__ pos, int i1, int a0, int a1, int lane,
                                              float radius_sq, int64_t out_base, const NlistOut& o) {
    const float p1[3] = {pos[3 * (int64_t)i1], pos[3 * (int64_t)i1 + 1], pos[3 * (int64_t)i1 + 2]};
    const double f1[3] = {
        (double)p1[0] * g.inv[0] + (double)p1[1] * g.inv[3] + (double)p1[2] * g.inv[6],
        (double)p1[0] * g.inv[1] + (double)p1[1] * g.inv[4] + (double)p1[2] * g.inv[7],
        (double)p1[0] * g.inv[2] + (double)p1[1] * g.inv[5] + (double)p1[2] * g.inv[8]};

    int total = 0;
    for (int chunk = a0; chunk < a1; chunk += 32) {
        const int i2 = chunk + lane;
        const bool valid = i2 < a1;
        float p2[3] = {0.f, 0.f, 0.f};
        int lo[3] = {1, 1, 1}, hi[3] = {0, 0, 0};
        if (valid) {
            p2[0] = pos[3 * (int64_t)i2]; p2[1] = pos[3 * (int64_t)i2 + 1]; p2[2] = pos[3 * (int64_t)i2 + 2];
            const double q0 = (double)p2[0], q1 = (double)p2[1], q2 = (double)p2[2];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double f2 = q0 * g.inv[k] + q1 * g.inv[3 + k] + q2 * g.inv[6 + k];
                double df = f1[k] - f2;
                double l = ceil(df - g.rho[k]), h = floor(df + g.rho[k]);
                lo[k] = (int)fmax(l, (double)-g.rep[k]);
                hi[k] = (int)fmin(h, (double)g.rep[k]);
            }
        }
        // pass A: count this lane's accepted images (ascending cell index = u1 slowest)
        int cnt = 0;
        float delta[3];
        for (int u1 = lo[0]; u1 <= hi[0]; ++u1)
            for (int u2 = lo[1]; u2 <= hi[1]; ++u2)
                for (int u3 = lo[2]; u3 <= hi[2]; ++u3) {
                    float d2 = exact_d2(g, p1, p2, u1, u2, u3, delta);
                    cnt += (d2 <= radius_sq && d2 > 0.0001f) ? 1 : 0;   // utils.py:202-205
                }
        // exclusive prefix over lanes -> position of this source's first edge inside the row
        int incl = cnt;
#pragma unroll
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, o2);
            if (lane >= o2) incl += v;
        }
        const int chunk_total = __shfl_sync(0xffffffffu, incl, 31);
        if (FILL && cnt > 0) {
            int64_t w = out_base + total + (incl - cnt);
            for (int u1 = lo[0]; u1 <= hi[0]; ++u1)
                for (int u2 = lo[1]; u2 <= hi[1]; ++u2)
                    for (int u3 = lo[2]; u3 <= hi[2]; ++u3) {
                        float d2 = exact_d2(g, p1, p2, u1, u2, u3, delta);
                        if (d2 <= radius_sq && d2 > 0.0001f) {
                            write_edge(o, w, i1, i2, u1, u2, u3, d2, delta);
                            ++w;
                        }
                    }
        }
        total += chunk_total;
    }
    return total;
}

template <bool FILL>
__global__ void __launch_bounds__(128)
nlist_kernel(const float* __restrict__ pos, const float* __restrict__ cell, const int32_t* __restrict__ crystal_ptr,
             const int32_t* __restrict__ node_crystal, int num_nodes, float radius, float radius_sq,
             const int32_t* __restrict__ reps, int reps_stride, int32_t* __restrict__ row_count,
             const int32_t* __restrict__ row_ptr, NlistOut o) {
    const int lane = threadIdx.x & 31;
    const int i1 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i1 >= num_nodes) return;
    const int b = node_crystal[i1];
    CellGeom g;
    load_geom(g, cell, reps, reps_stride, b, radius);
    const int total = scan_all_pairs<FILL>(g, pos, i1, crystal_ptr[b], crystal_ptr[b + 1], lane, radius_sq,
                                           FILL ? (int64_t)row_ptr[i1] : 0, o);
    if (!FILL && lane == 0) row_count[i1] = total;
}

// ------------------------------------------------------------------------------------------ cell list
// Large crystals: atoms are binned on a grid in (wrapped) fractional coordinates whose bins are at least rho_k wide
// (rho_k = the conservative image half-range of load_geom), so every (source, image) within the radius of a destination
// sits in one of the 27 bins around the destination's bin. A warp visits those bins, runs the reference's exact fp32 test
// on every candidate, and restores the reference's row order -- (source index, cell index) ascending, the flat
// (index2, cell) enumeration of utils.py:116-123,166-170 -- by a rank sort of the accepted keys in shared memory.
// Crystals that are too small for a useful grid (fewer than kMinBins bins) and rows longer than kRowCap keep the all-pairs scan.
constexpr int kMinBins = 64;
constexpr int kRowCap = 1024;

struct CrystalGrid {      // per crystal, in the workspace
    CellGeom g;           // computed once per crystal (fp64 inverse, image half-ranges): fp64 is slow on this part, and
                          // recomputing it in every warp was ~40 % of the XU pipe of the all-pairs kernels
    int32_t nb[3];        // bins per axis (0 = all-pairs path)
    int32_t bin0;         // first bin of this crystal in the global bin arrays
};
constexpr int kGridWords = (int)((sizeof(CrystalGrid) + 3) / 4);

__global__ void grid_setup_kernel(const float* __restrict__ cell, const int32_t* __restrict__ crystal_ptr, int num_crystals,
                                  float radius, const int32_t* __restrict__ reps, int reps_stride, CrystalGrid* __restrict__ grids) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num_crystals) return;
    CellGeom g;
    load_geom(g, cell, reps, reps_stride, b, radius);
    const int n = crystal_ptr[b + 1] - crystal_ptr[b];
    int nb[3];
    double tot = 1.0;
    for (int k = 0; k < 3; ++k) {
        nb[k] = g.rho[k] > 0.0 ? (int)floor(1.0 / g.rho[k]) : 1;      // bin width 1/nb >= rho
        if (nb[k] > 1024) nb[k] = 1024;
        tot *= (double)(nb[k] > 0 ? nb[k] : 0);
    }
    // never more bins than atoms (the bin arrays are sized by the atom count): coarser bins are always valid
    if (tot > (double)n && tot > 0.0) {
        const double sc = cbrt((double)n / tot);
        for (int k = 0; k < 3; ++k) { nb[k] = (int)floor(nb[k] * sc); if (nb[k] < 1) nb[k] = 1; }
    }
    const int64_t ncells = (int64_t)(2 * g.rep[0] + 1) * (2 * g.rep[1] + 1) * (2 * g.rep[2] + 1);
    const bool use = nb[0] >= 1 && nb[1] >= 1 && nb[2] >= 1 && (int64_t)nb[0] * nb[1] * nb[2] >= kMinBins && (int64_t)nb[0] * nb[1] * nb[2] <= n &&
                     ncells <= 1024 && n < (1 << 22);          // row keys pack (source, cell) into 22 + 10 bits
    grids[b].g = g;
    grids[b].nb[0] = use ? nb[0] : 0; grids[b].nb[1] = use ? nb[1] : 0; grids[b].nb[2] = use ? nb[2] : 0;
    grids[b].bin0 = crystal_ptr[b];          // bins of crystal b live at [crystal_ptr[b], crystal_ptr[b] + nb0*nb1*nb2) (<= its atom count)
}

// wrapped fractional coordinate, its bin and the integer shift that was removed (f = fw + shift)
__device__ __forceinline__ void frac_bin(const CellGeom& g, const float p[3], const int32_t nb[3], int bin[3], int shift[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double f = (double)p[0] * g.inv[k] + (double)p[1] * g.inv[3 + k] + (double)p[2] * g.inv[6 + k];
        const double fl = floor(f);
        shift[k] = (int)fl;
        int bk = (int)((f - fl) * (double)nb[k]);
        bin[k] = bk >= nb[k] ? nb[k] - 1 : (bk < 0 ? 0 : bk);
    }
}

// pass 0: bin id + shift of every atom, bin histogram
__global__ void bin_atoms_kernel(const float* __restrict__ pos, const float* __restrict__ cell, const int32_t* __restrict__ node_crystal,
                                 int num_nodes, float radius, const int32_t* __restrict__ reps, int reps_stride,
                                 const CrystalGrid* __restrict__ grids, int32_t* __restrict__ atom_bin, int32_t* __restrict__ bin_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_nodes) return;
    const int b = node_crystal[i];
    if (grids[b].nb[0] == 0) { atom_bin[i] = -1; return; }
    const CrystalGrid& gr = grids[b];
    const CellGeom& g = gr.g;
    const float p[3] = {pos[3 * (int64_t)i], pos[3 * (int64_t)i + 1], pos[3 * (int64_t)i + 2]};
    int bin[3], shift[3];
    frac_bin(g, p, gr.nb, bin, shift);
    const int id = gr.bin0 + (bin[0] * gr.nb[1] + bin[1]) * gr.nb[2] + bin[2];
    atom_bin[i] = id;
    atomicAdd(&bin_count[id], 1);            // integer atomics: order-independent result
}

__global__ void bin_place_kernel(const int32_t* __restrict__ atom_bin, int num_nodes, const int32_t* __restrict__ bin_ptr,
                                 int32_t* __restrict__ cursor, int32_t* __restrict__ bin_atoms) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_nodes) return;
    const int id = atom_bin[i];
    if (id < 0) return;
    bin_atoms[bin_ptr[id] + atomicAdd(&cursor[id], 1)] = i;      // order inside a bin is irrelevant: rows are sorted by key afterwards
}

template <bool FILL>
__global__ void __launch_bounds__(128)
nlist_cells_kernel(const float* __restrict__ pos, const float* __restrict__ cell, const int32_t* __restrict__ crystal_ptr,
                   const int32_t* __restrict__ node_crystal, int num_nodes, float radius, float radius_sq,
                   const int32_t* __restrict__ reps, int reps_stride, const CrystalGrid* __restrict__ grids,
                   const int32_t* __restrict__ bin_ptr, const int32_t* __restrict__ bin_atoms,
                   int32_t* __restrict__ row_count, const int32_t* __restrict__ row_ptr, NlistOut o) {
    __shared__ uint32_t keys_sm[FILL ? 4 * kRowCap : 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int i1 = blockIdx.x * (blockDim.x >> 5) + wib;
    if (i1 >= num_nodes) return;
    const int b = node_crystal[i1];
    const int a0 = crystal_ptr[b], a1 = crystal_ptr[b + 1];
    const CrystalGrid gr = grids[b];             // per-crystal geometry precomputed by grid_setup_kernel
    const CellGeom& g = gr.g;
    const int64_t out_base = FILL ? (int64_t)row_ptr[i1] : 0;
    const int row_len = FILL ? row_ptr[i1 + 1] - row_ptr[i1] : 0;
    if (gr.nb[0] == 0 || (FILL && row_len > kRowCap)) {
        const int total = scan_all_pairs<FILL>(g, pos, i1, a0, a1, lane, radius_sq, out_base, o);
        if (!FILL && lane == 0) row_count[i1] = total;
        return;
    }
    const float p1[3] = {pos[3 * (int64_t)i1], pos[3 * (int64_t)i1 + 1], pos[3 * (int64_t)i1 + 2]};
    int bin1[3], s1[3];
    frac_bin(g, p1, gr.nb, bin1, s1);
    const int ncell23 = (2 * g.rep[1] + 1) * (2 * g.rep[2] + 1), ncell3 = 2 * g.rep[2] + 1;
    uint32_t* keys = keys_sm + (FILL ? wib * kRowCap : 0);
    int total = 0;
    // 27 neighbour bins in the periodically extended grid: extended coordinate c = bin1 + d -> bin c mod nb, image v = floor(c / nb)
    for (int nbr = 0; nbr < 27; ++nbr) {
        int bb[3], v[3];
        const int d[3] = {nbr / 9 - 1, (nbr / 3) % 3 - 1, nbr % 3 - 1};
        // (fewer than 3 bins on an axis: the three extended coordinates still differ in v, nothing is visited twice)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int c = bin1[k] + d[k];
            v[k] = c < 0 ? -1 : (c >= gr.nb[k] ? 1 : 0);
            bb[k] = c - v[k] * gr.nb[k];
        }
        const int id = gr.bin0 + (bb[0] * gr.nb[1] + bb[1]) * gr.nb[2] + bb[2];
        const int q0 = bin_ptr[id], q1 = bin_ptr[id + 1];
        for (int q = q0 + lane; q < q1 + ((32 - ((q1 - q0) & 31)) & 31); q += 32) {       // whole warp iterates together (ballots below)
            bool ok = false;
            int i2 = 0, u[3] = {0, 0, 0};
            float d2 = 0.f, delta[3];
            if (q < q1) {
                i2 = bin_atoms[q];
                const float p2[3] = {pos[3 * (int64_t)i2], pos[3 * (int64_t)i2 + 1], pos[3 * (int64_t)i2 + 2]};
                int bin2[3], s2[3];
                frac_bin(g, p2, gr.nb, bin2, s2);
                // separation f1 - (f2 + u) = f1w - f2w - v  with  v = u + s2 - s1   =>   u = v - s2 + s1
                ok = true;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    u[k] = v[k] - s2[k] + s1[k];
                    ok = ok && u[k] >= -g.rep[k] && u[k] <= g.rep[k];      // the reference searches +-rep only (utils.py:166-170)
                }
                if (ok) {
                    d2 = exact_d2(g, p1, p2, u[0], u[1], u[2], delta);
                    ok = d2 <= radius_sq && d2 > 0.0001f;                  // utils.py:202-205
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (FILL && ok) {
                const int slot = total + __popc(m & ((1u << lane) - 1u));
                const uint32_t ci = (uint32_t)((u[0] + g.rep[0]) * ncell23 + (u[1] + g.rep[1]) * ncell3 + (u[2] + g.rep[2]));
                keys[slot] = ((uint32_t)(i2 - a0) << 10) | ci;             // (source, cell) key: needs atoms < 2^22 per crystal, cells < 1024
            }
            total += __popc(m);
        }
    }
    if (!FILL) {
        if (lane == 0) row_count[i1] = total;
        return;
    }
    __syncwarp();
    // rank sort of the row's keys (unique), then every lane writes the edges of its sorted positions
    uint32_t mine[kRowCap / 32];
    int rank[kRowCap / 32];
    const int per = (total + 31) / 32;
    for (int t = 0; t < per; ++t) {
        const int j = lane + 32 * t;
        mine[t] = j < total ? keys[j] : 0xffffffffu;
        rank[t] = 0;
    }
    for (int i = 0; i < total; ++i) {
        const uint32_t kv = keys[i];
        for (int t = 0; t < per; ++t) rank[t] += (kv < mine[t]) ? 1 : 0;
    }
    __syncwarp();
    for (int t = 0; t < per; ++t)
        if (lane + 32 * t < total) keys[rank[t]] = mine[t];
    __syncwarp();
    for (int w = lane; w < total; w += 32) {
        const uint32_t kv = keys[w];
        const int i2 = a0 + (int)(kv >> 10);
        int ci = (int)(kv & 1023u);
        const int u1 = ci / ncell23 - g.rep[0];
        ci %= ncell23;
        const int u2 = ci / ncell3 - g.rep[1], u3 = ci % ncell3 - g.rep[2];
        const float p2[3] = {pos[3 * (int64_t)i2], pos[3 * (int64_t)i2 + 1], pos[3 * (int64_t)i2 + 2]};
        float delta[3];
        const float d2 = exact_d2(g, p1, p2, u1, u2, u3, delta);
        write_edge(o, out_base + w, i1, i2, u1, u2, u3, d2, delta);
    }
}

__global__ void reps_kernel(const float* __restrict__ cell, int num_crystals, float radius, int pbc_mask,
                            int32_t* __restrict__ reps_out, int32_t* __restrict__ reps_max) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num_crystals) return;
    const float* c = cell + 9 * (int64_t)b;
    const float a1[3] = {c[0], c[1], c[2]}, a2[3] = {c[3], c[4], c[5]}, a3[3] = {c[6], c[7], c[8]};
    auto cross = [](const float* x, const float* y, float* o) {
        o[0] = __fsub_rn(__fmul_rn(x[1], y[2]), __fmul_rn(x[2], y[1]));
        o[1] = __fsub_rn(__fmul_rn(x[2], y[0]), __fmul_rn(x[0], y[2]));
        o[2] = __fsub_rn(__fmul_rn(x[0], y[1]), __fmul_rn(x[1], y[0]));
    };
    float c23[3], c31[3], c12[3];
    cross(a2, a3, c23);   // utils.py:135
    cross(a3, a1, c31);   // utils.py:145
    cross(a1, a2, c12);   // utils.py:152
    const float vol = __fadd_rn(__fadd_rn(__fmul_rn(a1[0], c23[0]), __fmul_rn(a1[1], c23[1])), __fmul_rn(a1[2], c23[2]));
    const float* cr[3] = {c23, c31, c12};
    for (int k = 0; k < 3; ++k) {
        int rep = 0;
        if (pbc_mask & (1 << k)) {
            float x = __fdiv_rn(cr[k][0], vol), y = __fdiv_rn(cr[k][1], vol), z = __fdiv_rn(cr[k][2], vol);
            float inv = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
            rep = (int)ceilf(__fmul_rn(radius, inv));   // utils.py:139-140
        }
        reps_out[3 * b + k] = rep;
        if (reps_max) atomicMax(&reps_max[k], rep);
    }
}

// Single-block exclusive scan: thread t owns a contiguous slice.
__global__ void __launch_bounds__(1024) scan_kernel(const int32_t* __restrict__ in, int n, int32_t* __restrict__ out) {
    __shared__ int32_t warp_tot[32];
    const int t = threadIdx.x, T = blockDim.x;
    const int per = (n + T - 1) / T;
    const int b0 = min(n, t * per), b1 = min(n, b0 + per);
    int32_t s = 0;
    for (int i = b0; i < b1; ++i) s += in[i];
    int32_t incl = s;
    const int lane = t & 31, w = t >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
        int32_t x = lane < (T >> 5) ? warp_tot[lane] : 0, xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t v = __shfl_up_sync(0xffffffffu, xi, o);
            if (lane >= o) xi += v;
        }
        warp_tot[lane] = xi - x;   // exclusive warp offsets
    }
    __syncthreads();
    int32_t run = warp_tot[w] + (incl - s);
    if (t == 0) out[0] = 0;
    for (int i = b0; i < b1; ++i) {
        run += in[i];
        out[i + 1] = run;
    }
}

__global__ void split_kernel(const int64_t* __restrict__ edge_index, int64_t E, int num_nodes,
                             int32_t* __restrict__ src32, int32_t* __restrict__ dst32, int32_t* __restrict__ flags) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= E) return;
    int64_t s = edge_index[e], d = edge_index[E + e];
    src32[e] = (int32_t)s;
    dst32[e] = (int32_t)d;
    if (e + 1 < E && edge_index[E + e + 1] < d) flags[0] = 1;                       // dst not sorted
    if (s < 0 || s >= num_nodes || d < 0 || d >= num_nodes) flags[1] = 1;           // out of range
}

__global__ void hist_kernel(const int32_t* __restrict__ keys, int64_t E, int32_t* __restrict__ counts) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e < E) atomicAdd(&counts[keys[e]], 1);   // integer atomics: order-independent result
}

__global__ void place_kernel(const int32_t* __restrict__ keys, int64_t E, const int32_t* __restrict__ ptr,
                             int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= E) return;
    int k = keys[e];
    int slot = atomicAdd(&cursor[k], 1);
    perm[ptr[k] + slot] = (int32_t)e;
}

// one warp per group: make the order inside each group ascending in edge id (=> deterministic perm)
__global__ void __launch_bounds__(128) group_sort_kernel(const int32_t* __restrict__ ptr, int num_nodes,
                                                         int32_t* __restrict__ perm) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= num_nodes) return;
    const int b0 = ptr[n], len = ptr[n + 1] - b0;
    if (len <= 1) return;
    int32_t* p = perm + b0;
    if (len <= 256) {   // rank sort, values held in registers
        int32_t val[8], rank[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int idx = lane + 32 * j;
            val[j] = idx < len ? p[idx] : 0x7fffffff;
            rank[j] = 0;
        }
        for (int i = 0; i < len; ++i) {
            int32_t v = p[i];
#pragma unroll
            for (int j = 0; j < 8; ++j) rank[j] += (v < val[j]) ? 1 : 0;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (lane + 32 * j < len) p[rank[j]] = val[j];
    } else {            // odd-even transposition sort in place (rare: in-degree > 256)
        for (int round = 0; round < len; ++round) {
            for (int i = (round & 1) + 2 * lane; i + 1 < len; i += 64) {
                int32_t a = p[i], c = p[i + 1];
                if (a > c) { p[i] = c; p[i + 1] = a; }
            }
            __syncwarp();
        }
    }
}

// kNN cap: one warp per destination row (see cartnet_nlist_knn_mask in the header)
__global__ void __launch_bounds__(128)
knn_mask_kernel(const float* __restrict__ direction, const int32_t* __restrict__ row_ptr, int num_nodes, int threshold,
                float tolerance, int strict, float* __restrict__ d2s, uint8_t* __restrict__ keep, int32_t* __restrict__ new_count) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= num_nodes) return;
    const int b0 = row_ptr[n], cnt = row_ptr[n + 1] - b0;
    if (cnt <= threshold) {          // cutoff would be inf (utils.py:294,322): everything stays
        for (int i = lane; i < cnt; i += 32) keep[b0 + i] = 1;
        if (lane == 0) new_count[n] = cnt;
        return;
    }
    for (int i = lane; i < cnt; i += 32) {
        const float* v = direction + 3 * (int64_t)(b0 + i);
        d2s[b0 + i] = __fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2]));   // utils.py:197
    }
    __syncwarp();
    const float* d2 = d2s + b0;
    // stable rank of every entry; the entry with rank == threshold carries the (threshold+1)-th smallest value
    float vk = 0.f;
    int kept = 0;
    for (int base = 0; base < cnt; base += 32) {
        const int i = base + lane;
        int rank = 0x7fffffff;
        float mine = 0.f;
        if (i < cnt) {
            mine = d2[i];
            rank = 0;
            for (int j = 0; j < cnt; ++j) {
                const float o = d2[j];
                rank += (o < mine || (o == mine && j < i)) ? 1 : 0;
            }
        }
        const unsigned hit = __ballot_sync(0xffffffffu, rank == threshold);
        if (hit) vk = __shfl_sync(0xffffffffu, mine, __ffs(hit) - 1);
        if (strict && i < cnt) {
            const int k = rank < threshold ? 1 : 0;
            keep[b0 + i] = (uint8_t)k;
            kept += k;
        }
    }
    if (!strict) {
        const float cutoff = __fadd_rn(vk, tolerance);            // utils.py:322-326
        for (int i = lane; i < cnt; i += 32) {
            const int k = d2[i] <= cutoff ? 1 : 0;
            keep[b0 + i] = (uint8_t)k;
            kept += k;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
    if (lane == 0) new_count[n] = kept;
}

}  // namespace cartnet

using namespace cartnet;

extern "C" {

int cartnet_nlist_reps(const float* cell, int32_t num_crystals, float radius, int32_t pbc_mask, int32_t* reps_out,
                       int32_t* reps_max, cartnet_stream_t stream) {
    CN_CHECK_ARG(cell && reps_out && num_crystals >= 0, "nlist_reps: bad arguments");
    if (num_crystals == 0) return 0;
    reps_kernel<<<ceil_div(num_crystals, 128), 128, 0, (cudaStream_t)stream>>>(cell, num_crystals, radius, pbc_mask,
                                                                                reps_out, reps_max);
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_nlist_count(const float* pos, const float* cell, const int32_t* crystal_ptr, const int32_t* node_crystal,
                        int32_t num_nodes, float radius, float radius_sq, const int32_t* reps, int32_t reps_stride,
                        int32_t* row_count, cartnet_stream_t stream) {
    CN_CHECK_ARG(pos && cell && crystal_ptr && node_crystal && reps && row_count, "nlist_count: null pointer");
    CN_CHECK_ARG(reps_stride == 0 || reps_stride == 3, "nlist_count: reps_stride must be 0 or 3");
    if (num_nodes <= 0) return 0;
    NlistOut o = {};
    nlist_kernel<false><<<ceil_div(num_nodes, 4), 128, 0, (cudaStream_t)stream>>>(
        pos, cell, crystal_ptr, node_crystal, num_nodes, radius, radius_sq, reps, reps_stride, row_count, nullptr, o);
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_exclusive_scan_i32(const int32_t* in, int32_t n, int32_t* out, cartnet_stream_t stream) {
    CN_CHECK_ARG(in && out && n >= 0, "exclusive_scan: bad arguments");
    scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(in, n, out);
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_nlist_fill(const float* pos, const float* cell, const int32_t* crystal_ptr, const int32_t* node_crystal,
                       int32_t num_nodes, float radius, float radius_sq, const int32_t* reps, int32_t reps_stride,
                       const int32_t* row_ptr, int64_t* edge_index, int64_t num_edges, float* unit_cell, float* dist,
                       float* direction, float* cart_dist, float* cart_dir, int32_t* src32, int32_t* dst32,
                       cartnet_stream_t stream) {
    CN_CHECK_ARG(pos && cell && crystal_ptr && node_crystal && reps && row_ptr, "nlist_fill: null pointer");
    CN_CHECK_ARG(reps_stride == 0 || reps_stride == 3, "nlist_fill: reps_stride must be 0 or 3");
    if (num_nodes <= 0 || num_edges <= 0) return 0;
    CN_CHECK_ARG(edge_index && unit_cell && dist && direction, "nlist_fill: null output");
    NlistOut o = {edge_index, num_edges, unit_cell, dist, direction, cart_dist, cart_dir, src32, dst32};
    nlist_kernel<true><<<ceil_div(num_nodes, 4), 128, 0, (cudaStream_t)stream>>>(
        pos, cell, crystal_ptr, node_crystal, num_nodes, radius, radius_sq, reps, reps_stride, nullptr, row_ptr, o);
    CN_LAUNCH_CHECK();
    return 0;
}

// workspace layout (int32 words): grids [kGridWords B] | atom_bin [N] | bin_count / cursor [N + 1] | bin_ptr [N + 2] | bin_atoms [N]
static inline int64_t cells_ws_words(int64_t N, int64_t B) { return (int64_t)kGridWords * B + 2 + N + (N + 1) + (N + 2) + N + 16; }

int64_t cartnet_nlist_cells_workspace(int32_t num_nodes, int32_t num_crystals) {
    return cells_ws_words(num_nodes, num_crystals) * (int64_t)sizeof(int32_t);
}

struct CellsWs {
    CrystalGrid* grids;
    int32_t *atom_bin, *bin_count, *bin_ptr, *bin_atoms;
};
static inline CellsWs cells_ws(void* ws, int64_t N, int64_t B) {
    int32_t* w = (int32_t*)ws;
    CellsWs c;
    c.grids = (CrystalGrid*)w; w += ((int64_t)kGridWords * B + 1) & ~(int64_t)1;      // keeps the arrays behind it 8-byte aligned
    c.atom_bin = w; w += N;
    c.bin_count = w; w += N + 1;
    c.bin_ptr = w; w += N + 2;
    c.bin_atoms = w;
    return c;
}

int cartnet_nlist_cells_build(const float* pos, const float* cell, const int32_t* crystal_ptr, const int32_t* node_crystal,
                              int32_t num_nodes, int32_t num_crystals, float radius, const int32_t* reps, int32_t reps_stride,
                              void* workspace, cartnet_stream_t stream) {
    CN_CHECK_ARG(pos && cell && crystal_ptr && node_crystal && reps && workspace, "nlist_cells_build: null pointer");
    CN_CHECK_ARG(reps_stride == 0 || reps_stride == 3, "nlist_cells_build: reps_stride must be 0 or 3");
    if (num_nodes <= 0 || num_crystals <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const CellsWs c = cells_ws(workspace, num_nodes, num_crystals);
    grid_setup_kernel<<<ceil_div(num_crystals, 128), 128, 0, st>>>(cell, crystal_ptr, num_crystals, radius, reps, reps_stride, c.grids);
    CN_LAUNCH_CHECK();
    CN_CUDA(cudaMemsetAsync(c.bin_count, 0, sizeof(int32_t) * (size_t)(num_nodes + 1), st));
    bin_atoms_kernel<<<ceil_div(num_nodes, 256), 256, 0, st>>>(pos, cell, node_crystal, num_nodes, radius, reps, reps_stride, c.grids,
                                                              c.atom_bin, c.bin_count);
    CN_LAUNCH_CHECK();
    scan_kernel<<<1, 1024, 0, st>>>(c.bin_count, num_nodes + 1, c.bin_ptr);
    CN_LAUNCH_CHECK();
    CN_CUDA(cudaMemsetAsync(c.bin_count, 0, sizeof(int32_t) * (size_t)(num_nodes + 1), st));
    bin_place_kernel<<<ceil_div(num_nodes, 256), 256, 0, st>>>(c.atom_bin, num_nodes, c.bin_ptr, c.bin_count, c.bin_atoms);
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_nlist_cells_count(const float* pos, const float* cell, const int32_t* crystal_ptr, const int32_t* node_crystal,
                              int32_t num_nodes, int32_t num_crystals, float radius, float radius_sq, const int32_t* reps,
                              int32_t reps_stride, const void* workspace, int32_t* row_count, cartnet_stream_t stream) {
    CN_CHECK_ARG(pos && cell && crystal_ptr && node_crystal && reps && workspace && row_count, "nlist_cells_count: null pointer");
    if (num_nodes <= 0) return 0;
    const CellsWs c = cells_ws(const_cast<void*>(workspace), num_nodes, num_crystals);
    NlistOut o = {};
    nlist_cells_kernel<false><<<ceil_div(num_nodes, 4), 128, 0, (cudaStream_t)stream>>>(
        pos, cell, crystal_ptr, node_crystal, num_nodes, radius, radius_sq, reps, reps_stride, c.grids, c.bin_ptr, c.bin_atoms,
        row_count, nullptr, o);
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_nlist_cells_fill(const float* pos, const float* cell, const int32_t* crystal_ptr, const int32_t* node_crystal,
                             int32_t num_nodes, int32_t num_crystals, float radius, float radius_sq, const int32_t* reps,
                             int32_t reps_stride, const void* workspace, const int32_t* row_ptr, int64_t* edge_index,
                             int64_t num_edges, float* unit_cell, float* dist, float* direction, float* cart_dist,
                             float* cart_dir, int32_t* src32, int32_t* dst32, cartnet_stream_t stream) {
    CN_CHECK_ARG(pos && cell && crystal_ptr && node_crystal && reps && workspace && row_ptr, "nlist_cells_fill: null pointer");
    if (num_nodes <= 0 || num_edges <= 0) return 0;
    CN_CHECK_ARG(edge_index && unit_cell && dist && direction, "nlist_cells_fill: null output");
    const CellsWs c = cells_ws(const_cast<void*>(workspace), num_nodes, num_crystals);
    NlistOut o = {edge_index, num_edges, unit_cell, dist, direction, cart_dist, cart_dir, src32, dst32};
    nlist_cells_kernel<true><<<ceil_div(num_nodes, 4), 128, 0, (cudaStream_t)stream>>>(
        pos, cell, crystal_ptr, node_crystal, num_nodes, radius, radius_sq, reps, reps_stride, c.grids, c.bin_ptr, c.bin_atoms,
        nullptr, row_ptr, o);
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_nlist_knn_mask(const float* direction, const int32_t* row_ptr, int32_t num_nodes, int32_t threshold, float tolerance,
                           int32_t strict, float* d2_scratch, uint8_t* keep, int32_t* new_row_count, cartnet_stream_t stream) {
    CN_CHECK_ARG(row_ptr && keep && new_row_count && d2_scratch && threshold > 0, "nlist_knn_mask: bad arguments");
    if (num_nodes <= 0) return 0;
    CN_CHECK_ARG(direction, "nlist_knn_mask: null direction");
    knn_mask_kernel<<<ceil_div(num_nodes, 4), 128, 0, (cudaStream_t)stream>>>(direction, row_ptr, num_nodes, threshold, tolerance,
                                                                             strict, d2_scratch, keep, new_row_count);
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_graph_split(const int64_t* edge_index, int64_t num_edges, int32_t num_nodes, int32_t* src32,
                        int32_t* dst32, int32_t* flags, cartnet_stream_t stream) {
    CN_CHECK_ARG(flags, "graph_split: null flags");
    CN_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(int32_t), (cudaStream_t)stream));
    if (num_edges <= 0) return 0;
    CN_CHECK_ARG(edge_index && src32 && dst32, "graph_split: null pointer");
    split_kernel<<<(unsigned)ceil_div64(num_edges, 256), 256, 0, (cudaStream_t)stream>>>(edge_index, num_edges,
                                                                                         num_nodes, src32, dst32, flags);
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_graph_csr(const int32_t* keys, int64_t num_edges, int32_t num_nodes, int32_t* ptr, int32_t* perm,
                      int32_t* cursor, cartnet_stream_t stream) {
    CN_CHECK_ARG(ptr && cursor && num_nodes >= 0, "graph_csr: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CN_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (size_t)num_nodes, st));
    if (num_edges > 0) {
        CN_CHECK_ARG(keys && perm, "graph_csr: null pointer");
        hist_kernel<<<(unsigned)ceil_div64(num_edges, 256), 256, 0, st>>>(keys, num_edges, cursor);
        CN_LAUNCH_CHECK();
    }
    scan_kernel<<<1, 1024, 0, st>>>(cursor, num_nodes, ptr);
    CN_LAUNCH_CHECK();
    if (num_edges > 0 && num_nodes > 0) {
        CN_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (size_t)num_nodes, st));
        place_kernel<<<(unsigned)ceil_div64(num_edges, 256), 256, 0, st>>>(keys, num_edges, ptr, cursor, perm);
        CN_LAUNCH_CHECK();
        group_sort_kernel<<<ceil_div(num_nodes, 4), 128, 0, st>>>(ptr, num_nodes, perm);
        CN_LAUNCH_CHECK();
    }
    return 0;
}

}  // extern "C"
