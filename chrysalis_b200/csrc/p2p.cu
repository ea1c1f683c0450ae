// One-shot all-gather of small per-gene statistics over NVLink / NVSwitch peer memory.
//
// The only exchange of the gene-sharded Moran path is the all-gather of the per-gene results (8 bytes per gene, 18 KB per
// rank at cfg4 on 8 GPUs).  A ring collective pays one hop latency per rank for it; here every rank WRITES its shard
// straight into the receive buffer of every peer (buffers of a symmetric allocation, mapped into every process), then raises
// a flag in each peer's signal pad and waits until all peers have raised theirs: one launch, latency independent of the
// number of ranks.  Ranks sit on different GPUs, so the kernels that wait for one another never share a device.
#include "common.cuh"

namespace chr {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// peer_bufs[p]: receive buffer of rank p (byte address), this rank's shard goes to element offset `dst_elem0`;
// signal_pads[p]: uint32 signal pad of rank p, slot `sig0 + rank` is raised to `epoch` by this rank
__global__ void __launch_bounds__(1024)
p2p_allgather_kernel(const double* __restrict__ local, int64_t count, void* const* __restrict__ peer_bufs, void* const* __restrict__ signal_pads,
                     int rank, int world, int64_t dst_elem0, int sig0, uint32_t epoch) {
    for (int p = 0; p < world; ++p) {
        double* dst = reinterpret_cast<double*>(peer_bufs[p]) + dst_elem0;
        for (int64_t i = threadIdx.x; i < count; i += blockDim.x) dst[i] = local[i];
    }
    __syncthreads();
    if ((int)threadIdx.x < world) {
        __threadfence_system();                                        // the shard written above is visible before the flag
        st_release_sys(reinterpret_cast<uint32_t*>(signal_pads[threadIdx.x]) + sig0 + rank, epoch);
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(signal_pads[rank]) + sig0 + threadIdx.x;
        for (uint32_t spin = 0; ld_acquire_sys(mine) != epoch; ++spin)
            if (spin > (1u << 28)) __trap();                           // a peer that never arrives becomes an error, not a hang
    }
    __syncthreads();
}

}  // namespace chr

extern "C" {

int chr_p2p_allgather_f64(const double* local, int64_t count, void* const* peer_bufs_dev, void* const* signal_pads_dev, int rank,
                          int world, int64_t dst_elem0, int signal_slot0, uint32_t epoch, chr_stream_t stream) {
    using namespace chr;
    CHR_REQUIRE(local && peer_bufs_dev && signal_pads_dev && count >= 0 && world > 0 && world <= 64 && rank >= 0 && rank < world && epoch != 0,
                "chr_p2p_allgather_f64: bad arguments");
    p2p_allgather_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(local, count, peer_bufs_dev, signal_pads_dev, rank, world, dst_elem0, signal_slot0, epoch);
    CHR_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
