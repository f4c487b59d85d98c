// Device-side batch assembly: replaces PyG's Batch.from_data_list / DataLoader collation on the host
// followed by batch.to("cuda:0") (/root/reference/loader/loader.py:114-124, train/train.py:169). The data set lives in
// HBM as per-field blobs, every blob the concatenation of all crystals; a batch is a list of crystal ids, and each
// output field is a segmented gather of the selected crystals' ranges -- with the batch's cumulative node / edge offset
// added where the field holds indices (what PyG does to edge_index), or replaced by the crystal's slot (the `batch`
// vector). One launch assembles every field of the batch, including the int32 CSR views the layer kernels use, so a
// training step never touches host memory for its inputs beyond the list of crystal ids.
#include "common.cuh"

namespace cartnet {

constexpr int kMaxFields = 32;

struct CollateTable {
    cartnet_collate_field_t f[kMaxFields];
    int64_t work_begin[kMaxFields + 1];      // prefix sum of elements over fields
    int num_fields;
};

// largest s with ptr[s] <= i (ptr ascending, ptr[0] = 0, ptr[n] = total > i)
__device__ __forceinline__ int upper_slot(const int32_t* __restrict__ ptr, int n, int64_t i) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)ptr[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
collate_kernel(const CollateTable t, const int32_t* __restrict__ ids, int num_sel,
               const int32_t* __restrict__ src_ptr /* [4][G+1] blob offsets per kind */, int64_t src_ptr_pitch,
               const int32_t* __restrict__ out_ptr /* [4][B+1] batch offsets per kind */) {
    const int64_t total = t.work_begin[t.num_fields];
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
        int fi = 0;
        while (fi + 1 < t.num_fields && w >= t.work_begin[fi + 1]) ++fi;
        const cartnet_collate_field_t f = t.f[fi];
        const int64_t i = w - t.work_begin[fi];                       // element of the output field
        const int32_t* op = out_ptr + (int64_t)f.kind * (num_sel + 1);
        int slot;
        int64_t local;
        if (f.kind == CARTNET_COLLATE_PER_GRAPH) { slot = (int)i; local = 0; }
        else { slot = upper_slot(op, num_sel, i); local = i - op[slot]; }
        const int g = ids[slot];
        const int64_t s = (int64_t)src_ptr[(int64_t)f.kind * src_ptr_pitch + g] + local;   // element of the blob
        const int32_t node_off = out_ptr[0 * (num_sel + 1) + slot], edge_off = out_ptr[1 * (num_sel + 1) + slot];
        switch (f.op) {
        case CARTNET_COLLATE_COPY: {
            if (f.elem_bytes == 1) { ((uint8_t*)f.dst)[i] = ((const uint8_t*)f.src)[s]; break; }
            const int nw = f.elem_bytes >> 2;
            const uint32_t* sp = (const uint32_t*)f.src + s * nw;
            uint32_t* dp = (uint32_t*)f.dst + i * nw;
            for (int k = 0; k < nw; ++k) dp[k] = sp[k];
            break;
        }
        case CARTNET_COLLATE_I32_PLUS_NODE: ((int32_t*)f.dst)[i] = ((const int32_t*)f.src)[s] + node_off; break;
        case CARTNET_COLLATE_I32_PLUS_EDGE: ((int32_t*)f.dst)[i] = ((const int32_t*)f.src)[s] + edge_off; break;
        case CARTNET_COLLATE_I32_TO_I64_PLUS_NODE: ((int64_t*)f.dst)[i] = (int64_t)((const int32_t*)f.src)[s] + node_off; break;
        case CARTNET_COLLATE_SLOT_I64: ((int64_t*)f.dst)[i] = (int64_t)slot; break;
        default: break;
        }
    }
}

// closing entries of the CSR pointers: ptr[N] = E
__global__ void collate_tail_kernel(int32_t* a, int32_t* b, int64_t n, int32_t value) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (a) a[n] = value;
        if (b) b[n] = value;
    }
}

}  // namespace cartnet

using namespace cartnet;

extern "C" {

int cartnet_collate(const cartnet_collate_field_t* fields /* host */, int32_t num_fields, const int32_t* ids, int32_t num_selected,
                    const int32_t* blob_ptr, int64_t blob_ptr_pitch, const int32_t* batch_ptr,
                    const int64_t* totals /* host [4]: nodes, edges, non-H atoms, graphs of the batch */, cartnet_stream_t stream) {
    CN_CHECK_ARG(fields && ids && blob_ptr && batch_ptr && totals && num_fields > 0 && num_fields <= kMaxFields && num_selected > 0,
                 "collate: bad arguments (at most %d fields)", kMaxFields);
    CollateTable t;
    t.num_fields = num_fields;
    int64_t acc = 0;
    for (int i = 0; i < num_fields; ++i) {
        const cartnet_collate_field_t& f = fields[i];
        CN_CHECK_ARG(f.kind >= 0 && f.kind <= 3 && f.dst && (f.src || f.op == CARTNET_COLLATE_SLOT_I64), "collate: field %d is malformed", i);
        CN_CHECK_ARG(f.op != CARTNET_COLLATE_COPY || f.elem_bytes == 1 || (f.elem_bytes > 0 && f.elem_bytes % 4 == 0),
                     "collate: field %d: element size must be 1 or a multiple of 4 bytes", i);
        t.f[i] = f;
        t.work_begin[i] = acc;
        acc += totals[f.kind];
    }
    t.work_begin[num_fields] = acc;
    if (acc == 0) return 0;
    int64_t blocks = ceil_div64(acc, 256);
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    collate_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(t, ids, num_selected, blob_ptr, blob_ptr_pitch, batch_ptr);
    CN_LAUNCH_CHECK();
    return 0;
}

int cartnet_collate_close_csr(int32_t* row_ptr, int32_t* col_ptr, int64_t num_nodes, int64_t num_edges, cartnet_stream_t stream) {
    CN_CHECK_ARG(num_edges < ((int64_t)1 << 31), "collate: more than 2^31 edges in one batch");
    collate_tail_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(row_ptr, col_ptr, num_nodes, (int32_t)num_edges);
    CN_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
