// Token all-gather of the batch-sharded decode as direct peer stores over NVLink: the one exchange of
// the data-parallel path is "every rank ends up with the (n_total, T+1) id matrix".  Instead of packing into a send
// buffer and calling a library collective on the compute stream, the rank that produced a shard's ids stores them
// (narrowed to int32) straight into slot `rank` of EVERY peer's receive buffer -- plain st.global to peer-mapped
// addresses, NVSwitch gives every pair full bandwidth -- then publishes one release flag per peer.  The reader
// kernel acquires the `world` flags of its own buffer and widens the rows into the caller's tensors.
//
//   receive buffer of a rank (symmetric: same layout everywhere, allocated by the host side as symmetric memory):
//     [2 parities][world slots][cap rows][T1 + 1] int32      cap = ceil(n_total / world); column T1 = the row's length
//     flags: [2 parities][world] uint32 sequence numbers + [2][world] int32 stop steps
//
// Pipelining contract (hmer-img2latex_b200/dist.py: TokenExchange): per step a rank runs  read(seq - 1)  BEFORE
// write(seq), both on its compute stream.  Two parities are then enough: a rank's write(seq + 2) is ordered after its
// read(seq + 1), which needs every peer's write(seq + 1), which that peer enqueued after its own read(seq) of the same
// parity.  The read of step i happens a whole step after the data was due, so the spin is normally satisfied at once
// and a slow rank does not stall the others' compute (the NCCL version made every step a barrier).
// The spin is bounded (~4 s of clock64): on expiry the kernel sets *timeout_flag and returns instead of hanging the GPU.
#include "common.cuh"

namespace i2l {
namespace {

constexpr int MAX_WORLD = 16;
struct Peers { int32_t* buf[MAX_WORLD]; uint32_t* flags[MAX_WORLD]; };

__host__ __device__ inline size_t slot_words(int cap, int T1) { return (size_t)cap * (T1 + 1); }

__global__ void xchg_write_kernel(const int64_t* __restrict__ tokens, const int32_t* __restrict__ lengths,
                                  const int32_t* __restrict__ steps, int b, int T1, int cap, int rank, int world,
                                  int parity, uint32_t seq, Peers P, unsigned int* counter) {
  const int W = T1 + 1;
  const size_t words = slot_words(cap, T1);
  const size_t off = ((size_t)parity * world + rank) * words;
  // one warp per row; a row is read once and stored to every peer (coalesced 128-byte segments)
  const int warps_per_block = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.x * warps_per_block + warp; row < cap; row += gridDim.x * warps_per_block) {
    for (int c = lane; c < W; c += 32) {
      int32_t v = -1;
      if (row < b) v = c < T1 ? (int32_t)tokens[(size_t)row * T1 + c] : lengths[row];
#pragma unroll 1
      for (int p = 0; p < world; ++p) P.buf[p][off + (size_t)row * W + c] = v;
    }
  }
  // all stores of this block are performed at system scope before the block is counted; the last block publishes
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x < (unsigned)world) {
      uint32_t* f = P.flags[threadIdx.x];
      reinterpret_cast<volatile int32_t*>(f)[2 * world + parity * world + rank] = *steps;   // stop step rides along
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f + parity * world + rank), "r"(seq) : "memory");
    }
    if (threadIdx.x == 0) *counter = 0;   // ready for the next launch (stream-ordered)
  }
}

__global__ void xchg_read_kernel(const int32_t* __restrict__ buf, const uint32_t* __restrict__ flags, int world, int cap,
                                 int T1, int n_total, int parity, uint32_t seq, int64_t* __restrict__ tokens,
                                 int32_t* __restrict__ lengths, int32_t* __restrict__ steps, int* timeout_flag) {
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  if (threadIdx.x < (unsigned)world) {
    const uint32_t* f = flags + parity * world + threadIdx.x;
    const long long t0 = clock64();
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (v == seq) break;
      if (clock64() - t0 > 8000000000LL) { ok = 0; break; }   // ~4 s at 1.9 GHz: report instead of hanging
      __nanosleep(200);
    } while (true);
  }
  __syncthreads();
  if (!ok) { if (threadIdx.x == 0) atomicExch(timeout_flag, 1); return; }
  const int W = T1 + 1;
  const size_t words = slot_words(cap, T1);
  const int base = n_total / world, rem = n_total % world;       // dist.shard_bounds: the first `rem` ranks hold one more
  const int warps_per_block = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g = blockIdx.x * warps_per_block + warp; g < n_total; g += gridDim.x * warps_per_block) {
    int r, local;
    if (g < rem * (base + 1)) { r = g / (base + 1); local = g - r * (base + 1); }
    else { r = rem + (g - rem * (base + 1)) / (base > 0 ? base : 1); local = g - rem * (base + 1) - (r - rem) * base; }
    const int32_t* src = buf + ((size_t)parity * world + r) * words + (size_t)local * W;
    for (int c = lane; c < T1; c += 32) tokens[(size_t)g * T1 + c] = (int64_t)src[c];
    if (lane == 0) lengths[g] = src[T1];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int m = 0;
    const int32_t* st = reinterpret_cast<const int32_t*>(flags) + 2 * world + parity * world;
    for (int r = 0; r < world; ++r) m = max(m, st[r]);          // the sticky stop composes as max over ranks
    *steps = m;
  }
}

}  // namespace
}  // namespace i2l

using namespace i2l;

extern "C" size_t i2l_token_exchange_buffer_bytes(int32_t world, int32_t n_total, int32_t T1) {
  if (world < 1 || world > MAX_WORLD || n_total < 1 || T1 < 1) return 0;
  const int cap = (n_total + world - 1) / world;
  return align_up(2 * (size_t)world * slot_words(cap, T1) * 4, 256) + 256 /* flags + steps */ + 256 /* counter */;
}

static size_t flags_offset(int world, int n_total, int T1) {
  const int cap = (n_total + world - 1) / world;
  return align_up(2 * (size_t)world * slot_words(cap, T1) * 4, 256);
}

extern "C" int i2l_token_exchange_write(const int64_t* tokens, const int32_t* lengths, const int32_t* steps, int32_t b,
                                        int32_t T1, int32_t n_total, int32_t rank, int32_t world,
                                        void* const* peer_buffers, uint32_t seq, void* stream) {
  I2L_TRY(device_check());
  I2L_REQUIRE(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, "token exchange: invalid rank / world");
  I2L_REQUIRE(steps && peer_buffers && b >= 0 && (b == 0 || (tokens && lengths)) && T1 >= 1 && n_total >= 1 && seq != 0,
              "token exchange: invalid arguments (seq must be non-zero)");
  const int cap = (n_total + world - 1) / world;
  I2L_REQUIRE(b <= cap, "token exchange: shard of %d rows exceeds the slot capacity %d", b, cap);
  Peers P{};
  const size_t fo = flags_offset(world, n_total, T1);
  for (int p = 0; p < world; ++p) {
    I2L_REQUIRE(peer_buffers[p] != nullptr, "token exchange: null peer buffer %d", p);
    P.buf[p] = reinterpret_cast<int32_t*>(peer_buffers[p]);
    P.flags[p] = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(peer_buffers[p]) + fo);
  }
  unsigned int* counter = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(peer_buffers[rank]) + fo + 256);
  // small footprint (<= 40 blocks of 16 warps): the pipelined exchange runs beside the persistent decode kernel, which
  // leaves 20 SMs idle; 5 MB of peer stores need a few us either way
  const int grid = min(cdiv(cap, 16), 40);
  KernelTimer kt("xchg.write_p2p", (cudaStream_t)stream);
  xchg_write_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>(tokens, lengths, steps, b, T1, cap, rank, world, (int)(seq & 1),
                                                            seq, P, counter);
  I2L_LAUNCH_OK();
  return I2L_OK;
}

extern "C" int i2l_token_exchange_read(const void* local_buffer, int32_t world, int32_t n_total, int32_t T1, uint32_t seq,
                                       int64_t* tokens, int32_t* lengths, int32_t* steps, int32_t* timeout_flag,
                                       void* stream) {
  I2L_TRY(device_check());
  I2L_REQUIRE(world >= 1 && world <= MAX_WORLD && local_buffer && tokens && lengths && steps && timeout_flag && seq != 0,
              "token exchange: invalid arguments");
  const int cap = (n_total + world - 1) / world;
  const char* base = reinterpret_cast<const char*>(local_buffer);
  const size_t fo = flags_offset(world, n_total, T1);
  const int grid = min(cdiv(n_total, 16), 40);
  KernelTimer kt("xchg.read_wait", (cudaStream_t)stream);
  xchg_read_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>(reinterpret_cast<const int32_t*>(base),
                                                           reinterpret_cast<const uint32_t*>(base + fo), world, cap, T1,
                                                           n_total, (int)(seq & 1), seq, tokens, lengths, steps, timeout_flag);
  I2L_LAUNCH_OK();
  return I2L_OK;
}
