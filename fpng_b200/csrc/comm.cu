// fpng_b200/csrc/comm.cu -- multi-GPU part of the C ABI: one process per GPU, images sharded across ranks with no
// collective on the data path, and ONE gather of the encoded (variable-size) files afterwards (BASELINE.json north_star,
// SURVEY.md section 8b/8e).  The reference has no distributed layer; everything here is new.
//
// The gather is a fused "compact + push" over NVLink peer memory, not a chain of NCCL point-to-point calls:
//   1. ncclAllGather of the per-file sizes (4 bytes per file; every rank learns every size, on the device),
//   2. gather_offsets_kernel: every rank derives the same table of 16-byte aligned byte offsets for all files of all ranks,
//   3. gather_push_kernel: every rank copies its own files with 128-bit stores STRAIGHT INTO THE RECEIVER'S WINDOW
//      (a cudaMalloc'ed buffer of the receiving process mapped here through CUDA IPC; NVLink 5 / NVSwitch peer stores),
//      at the offsets of step 2 -- no staging buffer, no host-side byte counts, no host synchronisation,
//   4. a 4-byte ncclAllReduce as the completion barrier: when it returns on the receiver, every sender's push kernel
//      has finished (stream order on the sender) and its stores are visible.
// dst_rank = -1 pushes to every rank (all-gather form: no single rank's NVLink ingress is the bottleneck).
// When peer windows cannot be mapped (no P2P / IPC), the same entry point falls back to grouped ncclSend/ncclRecv of the
// locally compacted shard, which needs ONE host synchronisation for the byte counts.
#include "../../include/fpng_b200.h"
#include "runtime.h"
#include <nccl.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <vector>

namespace fpngb {

// NCCL is bound at run time, on the first fpngb_comm_* call: a process that already holds an NCCL (a PyTorch process: its
// bundled libnccl.so.2) keeps using that one, others load the system library.  Linking libnccl directly would make THIS
// library pull the system NCCL into the process at load time and break a later `import torch` whose libtorch needs newer
// symbols from its own copy (same soname).
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*CommCount)(const ncclComm_t, int*);
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*);
    ncclResult_t (*CommCuDevice)(const ncclComm_t, int*);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    bool ok = false;
};
static NcclApi g_nccl;

static bool nccl_bind()
{
    if (g_nccl.ok) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);               // the copy the process already uses, if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
    bool all = true;
#define FPNGB_NCCL_SYM(field, name) do { *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) all = false; } while (0)
    FPNGB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); FPNGB_NCCL_SYM(CommInitRank, "ncclCommInitRank"); FPNGB_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    FPNGB_NCCL_SYM(CommCount, "ncclCommCount"); FPNGB_NCCL_SYM(CommUserRank, "ncclCommUserRank"); FPNGB_NCCL_SYM(CommCuDevice, "ncclCommCuDevice");
    FPNGB_NCCL_SYM(AllGather, "ncclAllGather"); FPNGB_NCCL_SYM(AllReduce, "ncclAllReduce"); FPNGB_NCCL_SYM(Send, "ncclSend"); FPNGB_NCCL_SYM(Recv, "ncclRecv");
    FPNGB_NCCL_SYM(GroupStart, "ncclGroupStart"); FPNGB_NCCL_SYM(GroupEnd, "ncclGroupEnd");
#undef FPNGB_NCCL_SYM
    g_nccl.ok = all;
    return all;
}
#define FPNGB_NEED_NCCL() do { if (!nccl_bind()) return FPNGB_ERR_INTERNAL; } while (0)

struct Comm {
    ncclComm_t nccl = nullptr;
    bool owned = false;
    int nranks = 0, rank = -1;
    // peer windows
    size_t window_bytes = 0;
    uint32_t nmax = 0;                           // file slots per rank
    uint8_t* window = nullptr;                   // this rank's receive window
    std::vector<uint8_t*> peer;                  // [nranks] mapped base of every rank's window (peer[rank] == window)
    bool p2p = false;
    uint8_t** d_peer = nullptr;                  // device copy of peer[]
    uint32_t* d_all_sizes = nullptr;             // [nranks * nmax]
    uint32_t* d_my_sizes = nullptr;              // [nmax]
    unsigned long long* d_offsets = nullptr;     // [nranks * nmax + 1]
    int* d_flag = nullptr;                       // completion barrier payload
    uint8_t* d_stage = nullptr; size_t stage_cap = 0;   // fallback path: locally compacted shard
};
static Comm g_comm;

#define FPNGB_NCCL_OK(expr) do { ncclResult_t r__ = (expr); if (r__ != ncclSuccess) return 2000 + (int)r__; } while (0)

// offsets[r * nmax + i] = byte offset of file i of rank r in the packed stream (rank order, 16-byte aligned files);
// offsets[nranks * nmax] = total.  One CTA; every rank computes the identical table.
__global__ void __launch_bounds__(1024) gather_offsets_kernel(const uint32_t* __restrict__ all_sizes, uint32_t total_slots,
                                                               unsigned long long* __restrict__ offsets)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < total_slots; i0 += blockDim.x) {
        const uint32_t i = i0 + tid;
        const unsigned long long v = i < total_slots ? (((unsigned long long)all_sizes[i] + 15ull) & ~15ull) : 0ull;
        unsigned long long s = v;
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, s, o); if (lane >= (uint32_t)o) s += u; }
        if (lane == 31) s_warp[warp] = s;
        __syncthreads();
        unsigned long long wb = 0, tot = 0;
        for (uint32_t k = 0; k < blockDim.x / 32; k++) { if (k < warp) wb += s_warp[k]; tot += s_warp[k]; }
        if (i < total_slots) offsets[i] = s_base + wb + s - v;
        __syncthreads();
        if (tid == 0) s_base += tot;
        __syncthreads();
    }
    if (tid == 0) offsets[total_slots] = s_base;
}

// grid (x, n_local, ndst): file blockIdx.y of this rank -> window of destination blockIdx.z (dst list in `dsts`)
__global__ void __launch_bounds__(256) gather_push_kernel(const uint8_t* __restrict__ files, size_t stride, const uint32_t* __restrict__ sizes,
                                                          unsigned long long* __restrict__ offsets, uint32_t my_slot0, uint32_t total_slots,
                                                          uint8_t* const* __restrict__ peer, int dst_rank, size_t cap)
{
    const uint32_t f = blockIdx.y;
    const int dst = dst_rank >= 0 ? dst_rank : (int)blockIdx.z;
    const uint32_t nvec = (sizes[f] + 15u) / 16u;
    const unsigned long long o = offsets[my_slot0 + f];
    if (o + (unsigned long long)nvec * 16ull > cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&offsets[total_slots], 1ull << 63);     // window too small: flagged, not copied
        return;
    }
    const uint4* src = reinterpret_cast<const uint4*>(files + (size_t)f * stride);
    uint4* d = reinterpret_cast<uint4*>(peer[dst] + o);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) d[i] = src[i];
}

__global__ void pad_sizes_kernel(const uint32_t* __restrict__ sizes, uint32_t n_local, uint32_t nmax, uint32_t* __restrict__ padded)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nmax) padded[i] = i < n_local ? sizes[i] : 0u;
}

static void comm_release_windows(Comm& c)
{
    for (int r = 0; r < (int)c.peer.size(); r++)
        if (c.peer[r] && r != c.rank) cudaIpcCloseMemHandle(c.peer[r]);
    c.peer.clear();
    if (c.window) cudaFree(c.window);
    if (c.d_peer) cudaFree(c.d_peer);
    if (c.d_all_sizes) cudaFree(c.d_all_sizes);
    if (c.d_my_sizes) cudaFree(c.d_my_sizes);
    if (c.d_offsets) cudaFree(c.d_offsets);
    if (c.d_flag) cudaFree(c.d_flag);
    if (c.d_stage) cudaFree(c.d_stage);
    c.window = nullptr; c.d_peer = nullptr; c.d_all_sizes = nullptr; c.d_my_sizes = nullptr; c.d_offsets = nullptr; c.d_flag = nullptr;
    c.d_stage = nullptr; c.stage_cap = 0; c.window_bytes = 0; c.nmax = 0; c.p2p = false;
}

}  // namespace fpngb

using namespace fpngb;

extern "C" {

int fpngb_comm_unique_id(void* id128)
{
    if (!id128) return FPNGB_ERR_INVALID_ARG;
    FPNGB_NEED_NCCL();
    static_assert(sizeof(ncclUniqueId) == FPNGB_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    FPNGB_NCCL_OK(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return FPNGB_OK;
}

int fpngb_comm_init(const void* id128, int nranks, int rank)
{
    Context& c = context();
    if (!c.ready) return FPNGB_ERR_NOT_INITIALIZED;
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return FPNGB_ERR_INVALID_ARG;
    FPNGB_NEED_NCCL();
    std::lock_guard<std::mutex> lk(c.mu);
    if (g_comm.nccl) return FPNGB_ERR_INVALID_ARG;            // one communicator per process; destroy it first
    FPNGB_CUDA_OK(cudaSetDevice(c.device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    FPNGB_NCCL_OK(g_nccl.CommInitRank(&g_comm.nccl, nranks, id, rank));
    g_comm.owned = true; g_comm.nranks = nranks; g_comm.rank = rank;
    return FPNGB_OK;
}

int fpngb_comm_adopt(void* nccl_comm)
{
    Context& c = context();
    if (!c.ready) return FPNGB_ERR_NOT_INITIALIZED;
    if (!nccl_comm) return FPNGB_ERR_INVALID_ARG;
    FPNGB_NEED_NCCL();
    std::lock_guard<std::mutex> lk(c.mu);
    if (g_comm.nccl) return FPNGB_ERR_INVALID_ARG;
    ncclComm_t comm = (ncclComm_t)nccl_comm;
    int n = 0, r = 0, dev = -1;
    FPNGB_NCCL_OK(g_nccl.CommCount(comm, &n));
    FPNGB_NCCL_OK(g_nccl.CommUserRank(comm, &r));
    FPNGB_NCCL_OK(g_nccl.CommCuDevice(comm, &dev));
    if (dev != c.device) return FPNGB_ERR_INVALID_ARG;
    g_comm.nccl = comm; g_comm.owned = false; g_comm.nranks = n; g_comm.rank = r;
    return FPNGB_OK;
}

int fpngb_comm_destroy(void)
{
    Context& c = context();
    if (!c.ready) return FPNGB_ERR_NOT_INITIALIZED;
    std::lock_guard<std::mutex> lk(c.mu);
    if (!g_comm.nccl) return FPNGB_OK;
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    comm_release_windows(g_comm);
    if (g_comm.owned) g_nccl.CommDestroy(g_comm.nccl);
    g_comm.nccl = nullptr; g_comm.nranks = 0; g_comm.rank = -1; g_comm.owned = false;
    return FPNGB_OK;
}

int fpngb_comm_info(int* nranks, int* rank, int* p2p)
{
    if (nranks) *nranks = g_comm.nranks;
    if (rank) *rank = g_comm.rank;
    if (p2p) *p2p = g_comm.p2p ? 1 : 0;
    return g_comm.nccl ? FPNGB_OK : FPNGB_ERR_NOT_INITIALIZED;
}

// Collective.  (Re)allocates this rank's receive window and per-gather tables and maps every peer's window.
int fpngb_gather_setup(size_t window_bytes, uint32_t max_files_per_rank)
{
    Context& c = context();
    if (!c.ready || !g_comm.nccl) return FPNGB_ERR_NOT_INITIALIZED;
    if (!window_bytes || !max_files_per_rank || max_files_per_rank > 65535u) return FPNGB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(c.mu);
    Comm& m = g_comm;
    FPNGB_CUDA_OK(cudaSetDevice(c.device));
    FPNGB_CUDA_OK(cudaDeviceSynchronize());
    comm_release_windows(m);
    window_bytes = align_up(window_bytes, 256);
    const uint32_t slots = (uint32_t)m.nranks * max_files_per_rank;
    FPNGB_CUDA_OK(cudaMalloc(&m.window, window_bytes));
    FPNGB_CUDA_OK(cudaMalloc(&m.d_peer, sizeof(uint8_t*) * m.nranks));
    FPNGB_CUDA_OK(cudaMalloc(&m.d_all_sizes, sizeof(uint32_t) * slots));
    FPNGB_CUDA_OK(cudaMalloc(&m.d_my_sizes, sizeof(uint32_t) * max_files_per_rank));
    FPNGB_CUDA_OK(cudaMalloc(&m.d_offsets, sizeof(unsigned long long) * (slots + 1)));
    FPNGB_CUDA_OK(cudaMalloc(&m.d_flag, 2 * sizeof(int)));
    FPNGB_CUDA_OK(cudaMemset(m.d_flag, 0, 2 * sizeof(int)));
    m.window_bytes = window_bytes; m.nmax = max_files_per_rank;

    // exchange CUDA IPC handles of the windows (64 bytes each) through the communicator itself
    cudaIpcMemHandle_t mine;
    const bool have_handle = cudaIpcGetMemHandle(&mine, m.window) == cudaSuccess;
    if (!have_handle) { memset(&mine, 0, sizeof mine); cudaGetLastError(); }
    struct Slot { cudaIpcMemHandle_t h; int ok; int pad[3]; };
    Slot my_slot; memset(&my_slot, 0, sizeof my_slot); my_slot.h = mine; my_slot.ok = have_handle ? 1 : 0;
    Slot* d_slots = nullptr;
    FPNGB_CUDA_OK(cudaMalloc(&d_slots, sizeof(Slot) * m.nranks));
    FPNGB_CUDA_OK(cudaMemcpy(d_slots + m.rank, &my_slot, sizeof my_slot, cudaMemcpyHostToDevice));
    cudaStream_t s = c.stream;
    ncclResult_t nr = g_nccl.AllGather(d_slots + m.rank, d_slots, sizeof(Slot), ncclChar, m.nccl, s);
    if (nr != ncclSuccess) { cudaFree(d_slots); return 2000 + (int)nr; }
    FPNGB_CUDA_OK(cudaStreamSynchronize(s));
    std::vector<Slot> all(m.nranks);
    FPNGB_CUDA_OK(cudaMemcpy(all.data(), d_slots, sizeof(Slot) * m.nranks, cudaMemcpyDeviceToHost));
    cudaFree(d_slots);
    m.peer.assign(m.nranks, nullptr);
    bool ok = true;
    for (int r = 0; r < m.nranks; r++) {
        if (r == m.rank) { m.peer[r] = m.window; continue; }
        void* p = nullptr;
        if (!all[r].ok || cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); continue; }
        m.peer[r] = (uint8_t*)p;
    }
    // every rank must take the same path: agree on min(ok)
    int h_ok = ok ? 1 : 0;
    FPNGB_CUDA_OK(cudaMemcpy(m.d_flag, &h_ok, sizeof(int), cudaMemcpyHostToDevice));
    FPNGB_NCCL_OK(g_nccl.AllReduce(m.d_flag, m.d_flag + 1, 1, ncclInt, ncclMin, m.nccl, s));
    FPNGB_CUDA_OK(cudaStreamSynchronize(s));
    FPNGB_CUDA_OK(cudaMemcpy(&h_ok, m.d_flag + 1, sizeof(int), cudaMemcpyDeviceToHost));
    m.p2p = h_ok == 1;
    if (m.p2p) FPNGB_CUDA_OK(cudaMemcpy(m.d_peer, m.peer.data(), sizeof(uint8_t*) * m.nranks, cudaMemcpyHostToDevice));
    else FPNGB_CUDA_OK(cudaMemcpy(m.d_peer + m.rank, &m.window, sizeof(uint8_t*), cudaMemcpyHostToDevice));
    return FPNGB_OK;
}

// Collective, enqueued on `stream`.  On return (host side) nothing has necessarily run yet: consume the outputs on the same stream.
// Receivers (dst_rank, or everybody for dst_rank = -1) find in *d_window the files of all ranks packed in rank order,
// (*d_offsets)[r * nmax + i] = byte offset of file i of rank r ((*d_offsets)[nranks * nmax] = total bytes, bit 63 = the window was too
// small), (*d_all_sizes)[r * nmax + i] = its size (0 for slots beyond a rank's file count).  nmax = max_files_per_rank of the setup.
int fpngb_gather_encoded_device(const void* d_files, size_t stride, const uint32_t* d_sizes, uint32_t n_local, int dst_rank,
                                void** d_window, uint64_t** d_offsets, uint32_t** d_all_sizes, void* stream)
{
    Context& c = context();
    Comm& m = g_comm;
    if (!c.ready || !m.nccl || !m.window) return FPNGB_ERR_NOT_INITIALIZED;
    if ((n_local && (!d_files || !d_sizes)) || n_local > m.nmax || dst_rank < -1 || dst_rank >= m.nranks) return FPNGB_ERR_INVALID_ARG;
    if (stride % 16 || (uintptr_t)d_files % 16) return FPNGB_ERR_ALIGNMENT;
    std::lock_guard<std::mutex> lk(c.mu);
    FPNGB_CUDA_OK(cudaSetDevice(c.device));
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t slots = (uint32_t)m.nranks * m.nmax;

    pad_sizes_kernel<<<(m.nmax + 255) / 256, 256, 0, s>>>(d_sizes, n_local, m.nmax, m.d_my_sizes);
    FPNGB_NCCL_OK(g_nccl.AllGather(m.d_my_sizes, m.d_all_sizes, m.nmax, ncclUint32, m.nccl, s));
    gather_offsets_kernel<<<1, 1024, 0, s>>>(m.d_all_sizes, slots, m.d_offsets);
    count_launch(2);
    if (m.p2p) {
        if (n_local) {
            const int ndst = dst_rank >= 0 ? 1 : m.nranks;
            dim3 grid(64, n_local, ndst);
            gather_push_kernel<<<grid, 256, 0, s>>>((const uint8_t*)d_files, stride, d_sizes, m.d_offsets, (uint32_t)m.rank * m.nmax, slots,
                                                    m.d_peer, dst_rank, m.window_bytes);
            count_launch(1);
        }
        FPNGB_CUDA_OK(cudaGetLastError());
        // completion barrier: returns on a receiver only after every sender's push kernel has completed
        FPNGB_NCCL_OK(g_nccl.AllReduce(m.d_flag, m.d_flag + 1, 1, ncclInt, ncclSum, m.nccl, s));
    } else {
        // fallback without peer windows: compact locally, then grouped send/recv with host-side byte counts (one synchronisation)
        std::vector<unsigned long long> h_off(slots + 1);
        FPNGB_CUDA_OK(cudaMemcpyAsync(h_off.data(), m.d_offsets, sizeof(unsigned long long) * (slots + 1), cudaMemcpyDeviceToHost, s));
        FPNGB_CUDA_OK(cudaStreamSynchronize(s));
        auto rank_begin = [&](int r) { return h_off[(size_t)r * m.nmax]; };
        auto rank_end = [&](int r) { return r + 1 < m.nranks ? h_off[(size_t)(r + 1) * m.nmax] : h_off[slots]; };
        const size_t my_bytes = (size_t)(rank_end(m.rank) - rank_begin(m.rank));
        if (h_off[slots] > m.window_bytes) { const unsigned long long flag = h_off[slots] | (1ull << 63);
            FPNGB_CUDA_OK(cudaMemcpyAsync(m.d_offsets + slots, &flag, sizeof flag, cudaMemcpyHostToDevice, s)); FPNGB_CUDA_OK(cudaStreamSynchronize(s)); }
        else {
            const bool i_receive = dst_rank < 0 || dst_rank == m.rank;
            uint8_t* my_dst = i_receive ? m.window + rank_begin(m.rank) : nullptr;
            if (!i_receive) {
                if (m.stage_cap < my_bytes) { if (m.d_stage) cudaFree(m.d_stage); m.d_stage = nullptr; m.stage_cap = 0;
                    FPNGB_CUDA_OK(cudaMalloc(&m.d_stage, align_up(my_bytes + 16, 1 << 20))); m.stage_cap = align_up(my_bytes + 16, 1 << 20); }
                my_dst = m.d_stage;
            }
            if (n_local) {
                // local compaction = the push kernel aimed at a local buffer whose base is shifted so that the rank's first file lands at 0
                uint8_t* shifted = my_dst - rank_begin(m.rank);
                FPNGB_CUDA_OK(cudaMemcpyAsync(m.d_peer + m.rank, &shifted, sizeof shifted, cudaMemcpyHostToDevice, s));
                dim3 grid(64, n_local, 1);
                gather_push_kernel<<<grid, 256, 0, s>>>((const uint8_t*)d_files, stride, d_sizes, m.d_offsets, (uint32_t)m.rank * m.nmax, slots,
                                                        m.d_peer, m.rank, (size_t)-1);
                count_launch(1);
            }
            FPNGB_NCCL_OK(g_nccl.GroupStart());
            for (int r = 0; r < m.nranks; r++) {
                if (r == m.rank) continue;
                const size_t rb = (size_t)(rank_end(r) - rank_begin(r));
                if (i_receive && rb) FPNGB_NCCL_OK(g_nccl.Recv(m.window + rank_begin(r), rb, ncclChar, r, m.nccl, s));
                if ((dst_rank < 0 || dst_rank == r) && my_bytes) FPNGB_NCCL_OK(g_nccl.Send(my_dst, my_bytes, ncclChar, r, m.nccl, s));
            }
            FPNGB_NCCL_OK(g_nccl.GroupEnd());
        }
    }
    FPNGB_CUDA_OK(cudaGetLastError());
    if (d_window) *d_window = m.window;
    if (d_offsets) *d_offsets = (uint64_t*)m.d_offsets;
    if (d_all_sizes) *d_all_sizes = m.d_all_sizes;
    return FPNGB_OK;
}

}  // extern "C"
