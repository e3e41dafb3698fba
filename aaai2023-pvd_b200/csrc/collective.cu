// collective.cu -- the one exchange step of the multi-GPU path: sum of the table gradient over the ranks.
//
// Rays shard across GPUs, parameters are replicated, so the only collective of a training step is the all-reduce of the hash
// table's gradient (42 MB fp32, 21 MB as the fp16 payload the reference accumulates these gradients in, gridencoder.cu:299-305).
// NCCL needs 67 us for those 21 MB on two B200 (scripts/micro/allreduce_probe.py) -- of the order of the whole compute step.
// On an NVSwitch system the reduction can be done BY THE SWITCH: the payload lives in a symmetric buffer that is also mapped
// through a multicast address; rank r owns 1/W of it, pulls the sum of all ranks' copies of its shard with
// multimem.ld_reduce (fp16 operands, fp32 accumulation in the switch) and pushes the result back to every rank with
// multimem.st.  Each link carries the shard once in and (W-1)/W... once out: two passes over 1/W of the data per GPU instead of
// NCCL's ring/tree steps.  MEASURED on 2 x B200 (scripts/micro/exchange_probe.py): this kernel 68 us for a 10.6 MB shard + two
// 13 us barriers + 11 us cast = 106 us, NCCL (cast + all-reduce) 89 us -- so NCCL stays the default and this path is opt-in
// (bench.py --grad-comm multimem) until it is tuned on more than two GPUs.  The barriers before (all payloads written) and after (all results visible) are the caller's
// (torch symmetric-memory signal pads: pvd_b200/dist.py).
#include "common.cuh"
#include "../../include/pvd_b200_fused.h"

namespace pvd {

__global__ void __launch_bounds__(256) k_multimem_allreduce_f16(__half* __restrict__ mc, uint64_t first_vec, uint64_t n_vec) {
    // four 16-byte switch reductions in flight per thread: the round trip through the NVSwitch is several microseconds
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_vec; i0 += 4 * stride) {
        uint32_t v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t i = i0 + u * stride;
            if (i < n_vec) {
                __half* p = mc + (first_vec + i) * 8;  // 8 halves = 16 bytes per lane
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0, %1, %2, %3}, [%4];"
                             : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3])
                             : "l"(p)
                             : "memory");
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t i = i0 + u * stride;
            if (i < n_vec) {
                __half* p = mc + (first_vec + i) * 8;
                asm volatile("multimem.st.relaxed.sys.global.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v[u][0]), "r"(v[u][1]),
                             "r"(v[u][2]), "r"(v[u][3])
                             : "memory");
            }
        }
    }
}

// ---- the same reduction as ONE launch: device-side barriers instead of two signal-pad barrier launches around the kernel ----------
// (each of those costs ~13 us on the stream; scripts/micro/exchange_probe.py).  Protocol per launch, all ranks run the same grid:
//   1. block 0, threads < W: cross-rank barrier over the symmetric-memory signal pads (CAS 0->1 on the peer's slot, CAS 1->0 on
//      mine: torch's own sync_remote_blocks protocol, other channel slots) -- every rank's payload is complete once its kernel
//      runs (stream order behind the cast); then `flag` releases this GPU's other blocks;
//   2. every block: multimem.ld_reduce of its part of this rank's shard, multimem.st of the sums to all ranks;
//   3. every block counts itself done; block 0 waits for all of them, fences at system scope, runs the cross-rank barrier again
//      (all ranks' stores into my copy are issued and fenced) and resets flag / counter for the next launch.
// Blocks only ever wait for block 0 (step 1) and block 0 only for blocks that wait for nothing else (step 3): no co-residency
// requirement beyond block 0 being scheduled, which a grid of at most 148 x 8 blocks of 256 threads guarantees.
// Every spin is bounded (~2 s of SM clocks): a peer that never arrives must not wedge the GPU; local[2] reports it.
constexpr long long kSpinLimit = 4000000000ll;
__device__ __forceinline__ bool sig_put(uint32_t* addr) {
    const long long t0 = clock64();
    while (atomicCAS_system(addr, 0u, 1u) != 0u)
        if (clock64() - t0 > kSpinLimit) return false;
    return true;
}
__device__ __forceinline__ bool sig_wait(uint32_t* addr) {
    const long long t0 = clock64();
    while (atomicCAS_system(addr, 1u, 0u) != 1u)
        if (clock64() - t0 > kSpinLimit) return false;
    return true;
}
__device__ __forceinline__ void cross_rank_barrier(uint32_t* const* pads, uint32_t rank, uint32_t world, uint32_t channel, uint32_t* err) {
    if (threadIdx.x < world) {
        const uint32_t peer = threadIdx.x;
        __threadfence_system();
        if (!sig_put(pads[peer] + channel * world + rank)) atomicExch(err, 1u);
        if (!sig_wait(pads[rank] + channel * world + peer)) atomicExch(err, 2u);
        __threadfence_system();
    }
}

// The same barrier with monotonic EPOCH flags instead of compare-and-swap pairs: rank r stores `epoch` into slot r of every peer's pad
// (a posted release store: no round trip) and polls its own W slots until all of them have reached `epoch` (acquire loads of LOCAL
// memory).  One NVLink one-way latency per barrier instead of the CAS protocol's round trips per peer -- measured on 8 GPUs: 33 us for
// the two barriers of one exchange with the CAS version (scripts/micro/exchange_probe.py).  Slots only ever grow; the epoch lives in
// the rank's persistent `local[3]`, advanced by the kernel itself (block 0), so launches need no host-side bookkeeping.
__device__ __forceinline__ void epoch_barrier(uint32_t* const* pads, uint32_t rank, uint32_t world, uint32_t channel, uint32_t epoch, uint32_t* err) {
    if (threadIdx.x < world) {
        const uint32_t peer = threadIdx.x;
        uint32_t* remote = pads[peer] + channel * world + rank;
        const uint32_t* mine = pads[rank] + channel * world + peer;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
        const long long t0 = clock64();
        for (;;) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((int32_t)(v - epoch) >= 0) break;
            if (clock64() - t0 > kSpinLimit) { atomicExch(err, 5u); break; }
        }
    }
}

template <int UNROLL>
__global__ void __launch_bounds__(256) k_multimem_allreduce_f16_fused(__half* __restrict__ mc, uint64_t first_vec, uint64_t n_vec,
                                                                      uint32_t* const* __restrict__ pads, uint32_t rank, uint32_t world,
                                                                      uint32_t* __restrict__ local /* [0] flag, [1] done, [2] error, [3] epoch */) {
    const uint32_t epoch0 = local[3];   // written only by block 0 at the very end of the previous launch: stable here
    if (blockIdx.x == 0) {
        epoch_barrier(pads, rank, world, 12u, epoch0 + 1u, local + 2);   // every rank's payload is written (its cast precedes its kernel)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicExch(local, 1u);
        }
    } else {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            while (atomicAdd(local, 0u) == 0u)
                if (clock64() - t0 > kSpinLimit) { atomicExch(local + 2, 3u); break; }
            __threadfence();
        }
        __syncthreads();
    }
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_vec; i0 += UNROLL * stride) {
        uint32_t v[UNROLL][4];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t i = i0 + u * stride;
            if (i < n_vec) {
                __half* p = mc + (first_vec + i) * 8;
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0, %1, %2, %3}, [%4];"
                             : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3])
                             : "l"(p)
                             : "memory");
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t i = i0 + u * stride;
            if (i < n_vec) {
                __half* p = mc + (first_vec + i) * 8;
                asm volatile("multimem.st.relaxed.sys.global.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v[u][0]), "r"(v[u][1]),
                             "r"(v[u][2]), "r"(v[u][3])
                             : "memory");
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        atomicAdd(local + 1, 1u);
    }
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            while (atomicAdd(local + 1, 0u) < gridDim.x)
                if (clock64() - t0 > kSpinLimit) { atomicExch(local + 2, 4u); break; }
            __threadfence_system();
        }
        __syncthreads();
        epoch_barrier(pads, rank, world, 12u, epoch0 + 2u, local + 2);   // every rank's multicast stores are issued and fenced
        __syncthreads();
        if (threadIdx.x == 0) {
            local[3] = epoch0 + 2u;
            local[1] = 0u;
            __threadfence();
            local[0] = 0u;
        }
    }
}

// ---- two-shot all-reduce over PEER pointers (no switch reduction): rank r owns 1/W of the payload, loads that shard from every
// rank's symmetric buffer with plain 16-byte NVLink loads (W independent loads in flight per thread), sums in fp32, and stores the
// fp16 result into every rank's buffer.  Per GPU: (W-1)/W of the payload in over NVLink and the same out -- both directions of the
// links are busy at once -- against multimem's fixed cost per switch request (scripts/micro/exchange_probe.py: 78 us for a 10.6 MB
// shard on two GPUs).  Barriers as above, inside the kernel.
template <int W, int UNROLL, bool WEAK>
__global__ void __launch_bounds__(512) k_p2p_allreduce_f16(__half* const* __restrict__ bufs, uint64_t first_vec, uint64_t n_vec,
                                                           uint32_t* const* __restrict__ pads, uint32_t rank,
                                                           uint32_t* __restrict__ local /* [0] flag, [1] done, [2] error */) {
    const uint32_t epoch0 = local[3];   // written only by block 0 at the very end of the previous launch: stable here
    if (blockIdx.x == 0) {
        epoch_barrier(pads, rank, W, 12u, epoch0 + 1u, local + 2);   // every rank's payload is written (its cast precedes its kernel)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicExch(local, 1u);
        }
    } else {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            while (atomicAdd(local, 0u) == 0u)
                if (clock64() - t0 > kSpinLimit) { atomicExch(local + 2, 3u); break; }
            __threadfence();
        }
        __syncthreads();
    }
    const uint4* src[W];
    uint4* dst[W];
#pragma unroll
    for (int r = 0; r < W; ++r) {
        // start with my own copy, then the peers in ring order: the ranks do not all hit the same peer at the same time
        const uint32_t q = (rank + (uint32_t)r) % (uint32_t)W;
        src[r] = reinterpret_cast<const uint4*>(bufs[q]) + first_vec;
        dst[r] = reinterpret_cast<uint4*>(bufs[q]) + first_vec;
    }
    // every address is read once per launch and written once: WEAK uses plain (non-coherent-path-free) accesses, ordered against the
    // other ranks by the system-scope fences of the two barriers; otherwise relaxed.sys accesses
    auto ld16 = [](const uint4* p) {
        uint4 v;
        if (WEAK) asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
        else asm volatile("ld.global.relaxed.sys.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
        return v;
    };
    auto st16 = [](uint4* p, const uint4& o) {
        if (WEAK) asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
        else asm volatile("st.global.relaxed.sys.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
    };
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_vec; i0 += UNROLL * stride) {
        uint4 v[UNROLL][W];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t i = i0 + u * stride;
#pragma unroll
            for (int r = 0; r < W; ++r) v[u][r] = (i < n_vec) ? ld16(src[r] + i) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t i = i0 + u * stride;
            if (i >= n_vec) break;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int r = 0; r < W; ++r) {
                const __half2* h = reinterpret_cast<const __half2*>(&v[u][r]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f = __half22float2(h[k]);
                    acc[2 * k] += f.x;
                    acc[2 * k + 1] += f.y;
                }
            }
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int k = 0; k < 4; ++k) oh[k] = __floats2half2_rn(acc[2 * k], acc[2 * k + 1]);
#pragma unroll
            for (int r = 0; r < W; ++r) st16(dst[r] + i, o);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        atomicAdd(local + 1, 1u);
    }
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            while (atomicAdd(local + 1, 0u) < gridDim.x)
                if (clock64() - t0 > kSpinLimit) { atomicExch(local + 2, 4u); break; }
            __threadfence_system();
        }
        __syncthreads();
        epoch_barrier(pads, rank, W, 12u, epoch0 + 2u, local + 2);   // every rank's stores into my copy are issued and fenced
        __syncthreads();
        if (threadIdx.x == 0) {
            local[3] = epoch0 + 2u;
            local[1] = 0u;
            __threadfence();
            local[0] = 0u;
        }
    }
}

// Link-rate probes for scripts/micro/exchange_probe.py: copy `n_vec` 16-byte vectors from the NEXT rank's buffer into mine (pull,
// mode 0), from mine into the next rank's (push, mode 1), or both at once on alternating vectors (mode 2).  No barriers: timing only.
__global__ void __launch_bounds__(512) k_p2p_copy_probe(__half* const* __restrict__ bufs, uint64_t n_vec, uint32_t rank, uint32_t world, uint32_t mode) {
    const uint4* peer_r = reinterpret_cast<const uint4*>(bufs[(rank + 1u) % world]);
    uint4* peer_w = reinterpret_cast<uint4*>(bufs[(rank + 1u) % world]);
    uint4* mine = reinterpret_cast<uint4*>(bufs[rank]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        if (mode == 0u) mine[i] = peer_r[i];
        else if (mode == 1u) peer_w[i] = mine[i];
        else if (i & 1u) mine[i] = peer_r[i];
        else peer_w[i] = mine[i];
    }
}

// fp32 gradient -> fp16 payload, and the inverse
__global__ void __launch_bounds__(256) k_f32_to_f16(const float4* __restrict__ src, uint2* __restrict__ dst, uint64_t n_vec4) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec4; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(src + i);
        const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
        dst[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
}

namespace {
template <int W>
void p2p_launch(uint32_t grid, cudaStream_t st, uint32_t unroll, bool weak, __half* const* bufs, uint64_t first_vec, uint64_t n_vec,
                       uint32_t* const* pads, uint32_t rank, uint32_t* local) {
    if (weak) {
        if (unroll >= 4) k_p2p_allreduce_f16<W, 4, true><<<grid, 512, 0, st>>>(bufs, first_vec, n_vec, pads, rank, local);
        else if (unroll >= 2) k_p2p_allreduce_f16<W, 2, true><<<grid, 512, 0, st>>>(bufs, first_vec, n_vec, pads, rank, local);
        else k_p2p_allreduce_f16<W, 1, true><<<grid, 512, 0, st>>>(bufs, first_vec, n_vec, pads, rank, local);
    } else {
        if (unroll >= 4) k_p2p_allreduce_f16<W, 4, false><<<grid, 512, 0, st>>>(bufs, first_vec, n_vec, pads, rank, local);
        else if (unroll >= 2) k_p2p_allreduce_f16<W, 2, false><<<grid, 512, 0, st>>>(bufs, first_vec, n_vec, pads, rank, local);
        else k_p2p_allreduce_f16<W, 1, false><<<grid, 512, 0, st>>>(bufs, first_vec, n_vec, pads, rank, local);
    }
}

}  // namespace

}  // namespace pvd

using namespace pvd;

extern "C" {

int pvd_multimem_allreduce_f16(void* multicast_ptr, uint64_t elem_offset, uint64_t elem_count, void* stream) {
    if (elem_count == 0) return PVD_OK;
    PVD_REQUIRE(multicast_ptr != nullptr && (elem_offset % 8u) == 0 && (elem_count % 8u) == 0);
    PVD_REQUIRE((reinterpret_cast<uintptr_t>(multicast_ptr) & 15u) == 0);
    const uint64_t n_vec = elem_count / 8u;
    const uint32_t grid = (uint32_t)min((unsigned long long)((n_vec + 1023u) / 1024u), 148ull * 8ull);
    k_multimem_allreduce_f16<<<grid, 256, 0, (cudaStream_t)stream>>>((__half*)multicast_ptr, elem_offset / 8u, n_vec);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_multimem_allreduce_f16_fused(void* multicast_ptr, uint64_t elem_offset, uint64_t elem_count, const void* signal_pad_ptrs_dev,
                                     uint32_t rank, uint32_t world, uint32_t* local_state, uint32_t blocks, uint32_t unroll, void* stream) {
    PVD_REQUIRE(multicast_ptr != nullptr && signal_pad_ptrs_dev != nullptr && local_state != nullptr);
    PVD_REQUIRE((elem_offset % 8u) == 0 && (elem_count % 8u) == 0 && world >= 1 && world <= 32 && rank < world);
    PVD_REQUIRE((reinterpret_cast<uintptr_t>(multicast_ptr) & 15u) == 0);
    const uint64_t n_vec = elem_count / 8u;
    uint32_t grid = blocks ? blocks : 64u;   // measured on 4 and 8 GPUs: 32-64 CTAs x unroll 2-4 (scripts/micro/exchange_probe.py)
    grid = max(1u, min(grid, 148u * 8u));
    uint32_t* const* pads = reinterpret_cast<uint32_t* const*>(signal_pad_ptrs_dev);
    cudaStream_t st = (cudaStream_t)stream;
    if (unroll >= 8) {
        k_multimem_allreduce_f16_fused<8><<<grid, 256, 0, st>>>((__half*)multicast_ptr, elem_offset / 8u, n_vec, pads, rank, world, local_state);
    } else if (unroll >= 4) {
        k_multimem_allreduce_f16_fused<4><<<grid, 256, 0, st>>>((__half*)multicast_ptr, elem_offset / 8u, n_vec, pads, rank, world, local_state);
    } else {
        k_multimem_allreduce_f16_fused<2><<<grid, 256, 0, st>>>((__half*)multicast_ptr, elem_offset / 8u, n_vec, pads, rank, world, local_state);
    }
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_p2p_allreduce_f16(const void* buffer_ptrs_dev, uint64_t elem_offset, uint64_t elem_count, const void* signal_pad_ptrs_dev,
                          uint32_t rank, uint32_t world, uint32_t* local_state, uint32_t blocks, uint32_t unroll, uint32_t weak,
                          void* stream) {
    PVD_REQUIRE(buffer_ptrs_dev != nullptr && signal_pad_ptrs_dev != nullptr && local_state != nullptr);
    PVD_REQUIRE((elem_offset % 8u) == 0 && (elem_count % 8u) == 0 && rank < world);
    const uint64_t n_vec = elem_count / 8u;
    uint32_t grid = blocks ? blocks : 148u;
    grid = max(1u, min(grid, 148u * 4u));
    __half* const* bufs = reinterpret_cast<__half* const*>(buffer_ptrs_dev);
    uint32_t* const* pads = reinterpret_cast<uint32_t* const*>(signal_pad_ptrs_dev);
    cudaStream_t st = (cudaStream_t)stream;
    switch (world) {
        case 2: p2p_launch<2>(grid, st, unroll, weak != 0, bufs, elem_offset / 8u, n_vec, pads, rank, local_state); break;
        case 4: p2p_launch<4>(grid, st, unroll, weak != 0, bufs, elem_offset / 8u, n_vec, pads, rank, local_state); break;
        case 8: p2p_launch<8>(grid, st, min(unroll, 2u), weak != 0, bufs, elem_offset / 8u, n_vec, pads, rank, local_state); break;
        default: return PVD_EUNSUPPORTED;
    }
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_p2p_copy_probe(const void* buffer_ptrs_dev, uint64_t elem_count, uint32_t rank, uint32_t world, uint32_t mode, uint32_t blocks, void* stream) {
    PVD_REQUIRE(buffer_ptrs_dev != nullptr && (elem_count % 8u) == 0 && rank < world && world >= 2);
    k_p2p_copy_probe<<<blocks ? blocks : 148u, 512, 0, (cudaStream_t)stream>>>(reinterpret_cast<__half* const*>(buffer_ptrs_dev), elem_count / 8u, rank,
                                                                              world, mode);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

int pvd_cast_f32_to_f16(const float* src, void* dst, uint64_t elem_count, void* stream) {
    if (elem_count == 0) return PVD_OK;
    PVD_REQUIRE(src && dst && (elem_count % 4u) == 0);
    PVD_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7u) == 0);
    const uint64_t n = elem_count / 4u;
    const uint32_t grid = (uint32_t)min((unsigned long long)((n + 255u) / 256u), 148ull * 16ull);
    k_f32_to_f16<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)src, (uint2*)dst, n);
    PVD_LAUNCH_CHECK();
    return PVD_OK;
}

}  // extern "C"
