// One-shot collectives over NVLink peer memory for the data-parallel step, sm_100a.
//
// The data-parallel CLEAR-VAE step exchanges only small tensors: the similarity operands + labels of the latent block
// (70 KB per rank at B = 1024, D = 8), its [B, 2] row statistics, the detached estimator latents (5 x 64 KB) and one
// 1.2 MB parameter-gradient vector.  At these sizes a ring collective is pure hop latency (7 hops at 8 GPUs), so each
// exchange is ONE kernel over peer-mapped buffers instead:
//
//   phase 1  every CTA copies this rank's pieces into the rank's own slot of its peer buffer, fences system-wide and
//            arrives on a local counter; the last CTA publishes `seq` to flag[rank] inside every peer's buffer
//            (st.release.sys over NVLink);
//   wait     `world` threads per CTA poll the local flags (relaxed system-scope loads, then one acquire fence) until every
//            peer has published `seq`;
//   phase 2  all threads PULL the peers' slots with volatile 16-byte loads over NVLink and write the final layout:
//            gather   -> dst[piece][rank][...] (each piece lands as one contiguous [world*B, ...] tensor: no cat / slice),
//            allreduce-> the element-wise sum in fixed rank order 0..world-1 (bit-identical on every rank) scattered back
//                        into the individual gradient tensors through a pointer table in kernel-parameter space.
//
// Two slots alternate with the call number: a rank can enter call k+2 only after every peer has published call k+1,
// i.e. after every peer has finished pulling call k, so slot (k & 1) is free again.  The call number lives on the device
// (advanced by the last CTA to finish), so launches are CUDA-graph replayable.  A poll that sees no progress for 20 s
// records an error in the buffer header and falls through instead of hanging the GPU.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kMaxR = CLEARVAE_PEER_MAX_RANKS;
constexpr int kMaxP = CLEARVAE_PEER_MAX_PIECES;
constexpr int kMaxT = CLEARVAE_ADAM_MAX_TENSORS;
constexpr int kHeader = CLEARVAE_PEER_HEADER_BYTES;
constexpr int kNT = 256;
constexpr unsigned long long kTimeoutNs = 20000000000ull;

struct PeerState {  // at byte 256 of the local buffer; flags[kMaxR] (written by the peers) sit at byte 0
  unsigned seq, arrive1, arrive2, error;
};

struct PeerTable {
  char* base[kMaxR];
  int world, rank;
  long long slot_bytes;
};

struct GatherArgs {
  const char* src[kMaxP];
  char* dst[kMaxP];
  unsigned bytes[kMaxP];  // per rank
  unsigned off[kMaxP];    // offset of the piece inside a slot (16-byte aligned)
  int vec[kMaxP];         // 1: 16-byte units, 0: 4-byte units
  unsigned ustart[kMaxP + 1];  // first flat unit of each piece
  int n;
};

struct ReduceArgs {
  float* t[kMaxT];
  int numel[kMaxT];
  int off[kMaxT + 1];  // float offsets inside a slot, multiples of 4
  int n;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_peer_v4(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_peer_u32(const void* p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// call number of this launch (every CTA reads it before the last finisher advances it)
__device__ __forceinline__ unsigned begin_call(PeerState* st, unsigned* s_seq) {
  if (threadIdx.x == 0) *s_seq = *reinterpret_cast<volatile unsigned*>(&st->seq) + 1u;
  __syncthreads();
  return *s_seq;
}

__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// publish phase-1 data to the peers, then wait until every peer has published the same call.
// Ordering chain: staging stores -> bar.sync -> thread 0: fence.gpu + arrive (gpu-scope atomic) -> last CTA observes the
// full count -> st.release.sys of the call number into every peer's flag word -> peer: relaxed polls + fence.acq_rel.sys
// -> bar.sync -> volatile pulls.  One gpu-scope fence per CTA and one system-scope release per peer, nothing per thread.
__device__ __forceinline__ void publish_and_wait(const PeerTable& pt, PeerState* st, unsigned seq, int* s_last) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    *s_last = (atomicAdd(&st->arrive1, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (*s_last && (int)threadIdx.x < pt.world)
    st_release_sys(reinterpret_cast<unsigned*>(pt.base[threadIdx.x]) + pt.rank, seq);
  if ((int)threadIdx.x < pt.world) {
    const unsigned* flag = reinterpret_cast<const unsigned*>(pt.base[pt.rank]) + threadIdx.x;
    const unsigned long long t0 = gtimer_ns();
    unsigned spins = 0;
    while ((int)(ld_relaxed_sys(flag) - seq) < 0) {
      if ((++spins & 1023u) == 0 && gtimer_ns() - t0 > kTimeoutNs) {
        atomicExch(&st->error, 1u + threadIdx.x);
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
}

__device__ __forceinline__ void end_call(PeerState* st, unsigned seq) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&st->arrive2, 1u) == gridDim.x - 1) {
      st->arrive1 = 0u;
      st->arrive2 = 0u;
      __threadfence();
      *reinterpret_cast<volatile unsigned*>(&st->seq) = seq;
    }
  }
}

// debug timeline: CTA 0 stamps %globaltimer at the phase boundaries of the last 64 calls (header bytes 512..3584)
__device__ __forceinline__ void stamp(const PeerTable& pt, unsigned seq, int k) {
  if (blockIdx.x == 0 && threadIdx.x == 0)
    reinterpret_cast<unsigned long long*>(pt.base[pt.rank] + 512)[(seq & 63u) * 6 + k] = gtimer_ns();
}

// piece that owns flat unit u (units of all pieces laid end to end)
__device__ __forceinline__ int piece_of(const GatherArgs& ga, unsigned u) {
  int k = 0;
#pragma unroll
  for (int j = 1; j < kMaxP; ++j)
    if (j < ga.n && u >= ga.ustart[j]) k = j;
  return k;
}

__global__ void __launch_bounds__(kNT) peer_gather_kernel(const __grid_constant__ PeerTable pt, const __grid_constant__ GatherArgs ga) {
  __shared__ unsigned s_seq;
  __shared__ int s_last;
  PeerState* st = reinterpret_cast<PeerState*>(pt.base[pt.rank] + 256);
  const unsigned seq = begin_call(st, &s_seq);
  stamp(pt, seq, 0);
  const long long slot = kHeader + (long long)(seq & 1u) * pt.slot_bytes;
  const unsigned gtid = blockIdx.x * kNT + threadIdx.x, gthreads = gridDim.x * kNT;
  char* mine = pt.base[pt.rank] + slot;
  const unsigned T = ga.ustart[ga.n];
  // phase 1: stage this rank's pieces (flat over all pieces: one load -> store chain per thread, not one per piece)
  for (unsigned u = gtid; u < T; u += gthreads) {
    const int k = piece_of(ga, u);
    const unsigned i = u - ga.ustart[k];
    if (ga.vec[k]) reinterpret_cast<uint4*>(mine + ga.off[k])[i] = reinterpret_cast<const uint4*>(ga.src[k])[i];
    else reinterpret_cast<unsigned*>(mine + ga.off[k])[i] = reinterpret_cast<const unsigned*>(ga.src[k])[i];
  }
  stamp(pt, seq, 1);
  publish_and_wait(pt, st, seq, &s_last);
  stamp(pt, seq, 2);
  // phase 2: pull every rank's slot; kU independent NVLink loads in flight per thread
  constexpr int kU = 4;
  const unsigned total = T * (unsigned)pt.world;
  for (unsigned base = gtid; base < total; base += gthreads * kU) {
    uint4 v[kU];
    char* d[kU];
    bool vec[kU], on[kU];
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      const unsigned idx = base + j * gthreads;
      on[j] = idx < total;
      vec[j] = false;
      d[j] = nullptr;
      if (on[j]) {
        const unsigned r = idx / T, u = idx - r * T;
        const int k = piece_of(ga, u);
        const unsigned i = u - ga.ustart[k];
        vec[j] = ga.vec[k] != 0;
        const unsigned unit = vec[j] ? 16u : 4u;
        const char* s = pt.base[r] + slot + ga.off[k] + (size_t)i * unit;
        d[j] = ga.dst[k] + (size_t)r * ga.bytes[k] + (size_t)i * unit;
        if (vec[j]) v[j] = ld_peer_v4(s);
        else v[j].x = ld_peer_u32(s);
      }
    }
#pragma unroll
    for (int j = 0; j < kU; ++j)
      if (on[j]) {
        if (vec[j]) *reinterpret_cast<uint4*>(d[j]) = v[j];
        else *reinterpret_cast<unsigned*>(d[j]) = v[j].x;
      }
  }
  stamp(pt, seq, 3);
  end_call(st, seq);
}

// tensor that owns float offset f of the packed vector: last k with off[k] <= f
__device__ __forceinline__ int tensor_of(const int* s_off, int n, int f) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (s_off[mid] <= f) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(kNT) peer_allreduce_kernel(const __grid_constant__ PeerTable pt, const __grid_constant__ ReduceArgs ra) {
  __shared__ unsigned s_seq;
  __shared__ int s_last;
  __shared__ int s_off[kMaxT + 1];
  PeerState* st = reinterpret_cast<PeerState*>(pt.base[pt.rank] + 256);
  for (int i = threadIdx.x; i <= ra.n; i += kNT) s_off[i] = ra.off[i];
  const unsigned seq = begin_call(st, &s_seq);   // (contains the barrier that publishes s_off)
  stamp(pt, seq, 0);
  const long long slot = kHeader + (long long)(seq & 1u) * pt.slot_bytes;
  const unsigned gtid = blockIdx.x * kNT + threadIdx.x, gthreads = gridDim.x * kNT;
  float4* mine = reinterpret_cast<float4*>(pt.base[pt.rank] + slot);
  const unsigned nvec = (unsigned)ra.off[ra.n] / 4u;
  // phase 1: pack (flat over the padded vector: offsets are multiples of 4, so a float4 never straddles two tensors)
  for (unsigned v = gtid; v < nvec; v += gthreads) {
    const int f = (int)(v * 4u);
    const int k = tensor_of(s_off, ra.n, f);
    const int i = f - s_off[k], n = ra.numel[k];
    const float* __restrict__ s = ra.t[k] + i;
    float4 x;
    if (i + 3 < n && (reinterpret_cast<uintptr_t>(s) & 15u) == 0) {
      x = *reinterpret_cast<const float4*>(s);
    } else {
      x.x = i < n ? s[0] : 0.f;
      x.y = i + 1 < n ? s[1] : 0.f;
      x.z = i + 2 < n ? s[2] : 0.f;
      x.w = i + 3 < n ? s[3] : 0.f;
    }
    mine[v] = x;
  }
  stamp(pt, seq, 1);
  publish_and_wait(pt, st, seq, &s_last);
  stamp(pt, seq, 2);
  // phase 2: pull + sum in rank order (world x kU independent NVLink loads in flight per thread) + scatter
  constexpr int kU = 2;
  for (unsigned base = gtid; base < nvec; base += gthreads * kU) {
    uint4 x[kU][kMaxR];
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      const unsigned v = base + j * gthreads;
#pragma unroll
      for (int r = 0; r < kMaxR; ++r)
        if (r < pt.world && v < nvec) x[j][r] = ld_peer_v4(pt.base[r] + slot + (size_t)v * 16);
    }
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      const unsigned v = base + j * gthreads;
      if (v >= nvec) continue;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < kMaxR; ++r)
        if (r < pt.world) {
          acc.x += __uint_as_float(x[j][r].x);
          acc.y += __uint_as_float(x[j][r].y);
          acc.z += __uint_as_float(x[j][r].z);
          acc.w += __uint_as_float(x[j][r].w);
        }
      const int f = (int)(v * 4u);
      const int k = tensor_of(s_off, ra.n, f);
      const int i = f - s_off[k], n = ra.numel[k];
      float* d = ra.t[k] + i;
      if (i + 3 < n && (reinterpret_cast<uintptr_t>(d) & 15u) == 0) {
        *reinterpret_cast<float4*>(d) = acc;
      } else {
        if (i < n) d[0] = acc.x;
        if (i + 1 < n) d[1] = acc.y;
        if (i + 2 < n) d[2] = acc.z;
        if (i + 3 < n) d[3] = acc.w;
      }
    }
  }
  stamp(pt, seq, 3);
  end_call(st, seq);
}

int fill_table(PeerTable& pt, void* const* bases_host, int32_t world, int32_t rank, int64_t buffer_bytes) {
  if (!bases_host || world < 1 || world > kMaxR || rank < 0 || rank >= world || buffer_bytes < kHeader + 32) return CLEARVAE_EINVAL;
  for (int r = 0; r < world; ++r) {
    if (!bases_host[r]) return CLEARVAE_EINVAL;
    pt.base[r] = static_cast<char*>(bases_host[r]);
  }
  pt.world = world;
  pt.rank = rank;
  pt.slot_bytes = ((buffer_bytes - kHeader) / 2) & ~15ll;
  return 0;
}

}  // namespace

extern "C" {

int clearvae_peer_alloc(int64_t bytes, void** ptr) {
  if (!ptr || bytes < kHeader + 32) return CLEARVAE_EINVAL;
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(*ptr, 0, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaDeviceSynchronize();
}

int clearvae_peer_free(void* ptr) { return ptr ? (int)cudaFree(ptr) : CLEARVAE_EINVAL; }

int clearvae_peer_export(void* ptr, uint8_t* handle_host) {
  if (!ptr || !handle_host) return CLEARVAE_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == CLEARVAE_PEER_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) return (int)e;
  memcpy(handle_host, &h, sizeof(h));
  return 0;
}

int clearvae_peer_open(const uint8_t* handle_host, void** ptr) {
  if (!ptr || !handle_host) return CLEARVAE_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

int clearvae_peer_close(void* ptr) { return ptr ? (int)cudaIpcCloseMemHandle(ptr) : CLEARVAE_EINVAL; }

int clearvae_peer_error(const void* local_base, int32_t* err_host) {
  if (!local_base || !err_host) return CLEARVAE_EINVAL;
  unsigned v = 0;
  cudaError_t e = cudaMemcpy(&v, static_cast<const char*>(local_base) + 256 + offsetof(PeerState, error), sizeof(v), cudaMemcpyDeviceToHost);
  *err_host = (int32_t)v;
  return (int)e;
}

int clearvae_peer_timeline(const void* local_base, uint64_t* stamps_host) {
  if (!local_base || !stamps_host) return CLEARVAE_EINVAL;
  return (int)cudaMemcpy(stamps_host, static_cast<const char*>(local_base) + 512, 64 * 6 * sizeof(uint64_t), cudaMemcpyDeviceToHost);
}

int clearvae_peer_gather(void* const* bases_host, int32_t world, int32_t rank, int64_t buffer_bytes, int32_t n_pieces,
                         const void* const* src_host, void* const* dst_host, const int64_t* bytes_host, void* stream) {
  PeerTable pt{};
  int rc = fill_table(pt, bases_host, world, rank, buffer_bytes);
  if (rc) return rc;
  if (n_pieces < 1 || n_pieces > kMaxP || !src_host || !dst_host || !bytes_host) return CLEARVAE_EINVAL;
  GatherArgs ga{};
  long long off = 0, units = 0;
  for (int k = 0; k < n_pieces; ++k) {
    if (!src_host[k] || !dst_host[k] || bytes_host[k] <= 0 || (bytes_host[k] & 3) || bytes_host[k] > 0x7fffffffLL) return CLEARVAE_EINVAL;
    ga.src[k] = static_cast<const char*>(src_host[k]);
    ga.dst[k] = static_cast<char*>(dst_host[k]);
    ga.bytes[k] = (unsigned)bytes_host[k];
    ga.off[k] = (unsigned)off;
    ga.vec[k] = ((bytes_host[k] & 15) == 0 && (reinterpret_cast<uintptr_t>(src_host[k]) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(dst_host[k]) & 15) == 0) ? 1 : 0;
    off += (bytes_host[k] + 15) & ~15ll;
    ga.ustart[k] = (unsigned)units;
    units += bytes_host[k] / (ga.vec[k] ? 16 : 4);
    if (units > 0x3fffffffLL) return CLEARVAE_EINVAL;
  }
  ga.ustart[n_pieces] = (unsigned)units;
  if (off > pt.slot_bytes) return CLEARVAE_EWORKSPACE;
  ga.n = n_pieces;
  const long long work = units * world;  // pull phase dominates
  const int grid = (int)std::min<long long>(128, std::max<long long>(1, (work + kNT * 2 - 1) / (kNT * 2)));
  peer_gather_kernel<<<grid, kNT, 0, (cudaStream_t)stream>>>(pt, ga);
  CV_LAUNCH_CHECK();
  return 0;
}

int clearvae_peer_allreduce(void* const* bases_host, int32_t world, int32_t rank, int64_t buffer_bytes, int32_t n_tensors,
                            float* const* tensors_host, const int64_t* numel_host, void* stream) {
  PeerTable pt{};
  int rc = fill_table(pt, bases_host, world, rank, buffer_bytes);
  if (rc) return rc;
  if (n_tensors <= 0) return 0;
  if (!tensors_host || !numel_host) return CLEARVAE_EINVAL;
  for (int t0 = 0; t0 < n_tensors; t0 += kMaxT) {
    ReduceArgs ra{};
    const int n = std::min(kMaxT, n_tensors - t0);
    long long off = 0;
    for (int i = 0; i < n; ++i) {
      const int j = t0 + i;
      if (!tensors_host[j] || numel_host[j] <= 0 || numel_host[j] > 0x3fffffffLL) return CLEARVAE_EINVAL;
      ra.t[i] = tensors_host[j];
      ra.numel[i] = (int)numel_host[j];
      ra.off[i] = (int)off;
      off += (numel_host[j] + 3) & ~3ll;
      if (off > 0x7ffffff0LL) return CLEARVAE_EINVAL;
    }
    ra.off[n] = (int)off;
    ra.n = n;
    if (off * 4 > pt.slot_bytes) return CLEARVAE_EWORKSPACE;
    const long long nvec = off / 4;
    const int grid = (int)std::min<long long>(256, std::max<long long>(1, (nvec + kNT * 2 - 1) / (kNT * 2)));
    peer_allreduce_kernel<<<grid, kNT, 0, (cudaStream_t)stream>>>(pt, ra);
    CV_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
