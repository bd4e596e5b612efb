// dx_bits.cuh -- warp-cooperative MSB-first bit writer shared by the QV encoder and the 2-bit packer.
//
// A warp produces one bit string: in every round each lane contributes a (possibly empty) run of
// bits that follows its lower neighbour's.  Lanes shift their pieces through a 64-bit register and
// store the 32-bit words they COMPLETE into the warp's staging area in shared memory; the partial
// words between neighbouring lanes are merged with a segmented OR-scan over shuffles, and the
// warp's trailing partial word travels in a register.  No shared-memory atomics.
#pragma once

#include "dx_common.cuh"

// Everything one warp needs to walk one line in rows of 32 chunks x 16 bytes.
struct LineWalk
{ const uint8_t *base;      // 16-byte aligned address at or before the line
  int32_t skew, rlen, nchunk;
  __device__ __forceinline__ void set(const uint8_t *line, int32_t len)
  { skew = (int32_t) (reinterpret_cast<uintptr_t>(line) & 15);
    base = line - skew; rlen = len; nchunk = (skew + len + 15) >> 4;
  }
  __device__ __forceinline__ uint32_t valid(int32_t c) const
  { if (c >= nchunk) return 0;
    const int32_t p0 = c*16 - skew;
    return dx_range16(max(0,-p0),min(16,rlen - p0));
  }
};

// staged words -> global at any byte alignment; SWAP: the stage holds MSB-first words that are
// to appear in the file as a byte string (the 2-bit packed tags), so every word is byte-swapped
template <bool SWAP>
__device__ __forceinline__ void copy_out(uint8_t *gdst, const uint32_t *ssrc, uint32_t n, int lane)
{ const uint8_t *sb = reinterpret_cast<const uint8_t *>(ssrc);
  uint32_t head = (4u - (uint32_t) (reinterpret_cast<uintptr_t>(gdst) & 3u)) & 3u;
  if (head > n) head = n;
  if ((uint32_t) lane < head)
    gdst[lane] = sb[SWAP ? (lane ^ 3) : lane];
  const uint32_t body = (n - head) >> 2;
  uint32_t *gw = reinterpret_cast<uint32_t *>(gdst + head);
  const uint32_t sh = head * 8u;                   // source is `head` bytes ahead of a word
  for (uint32_t i = lane; i < body; i += 32)
    { uint32_t lo = ssrc[i], hi = ssrc[i+1];
      if (SWAP) { lo = __byte_perm(lo,0,0x0123); hi = __byte_perm(hi,0,0x0123); }
      gw[i] = __funnelshift_r(lo,hi,sh);          // sh == 0 -> lo
    }
  const uint32_t done = head + 4u*body;
  if ((uint32_t) lane < n - done)
    gdst[done + lane] = sb[SWAP ? ((done + lane) ^ 3u) : (done + lane)];
}

// ---- the warp's output: completed words in shared memory, the trailing partial word in a register
struct WarpBits
{ uint32_t *stage;          // [kStageWords + 4]
  uint32_t  nst;            // completed words staged                        (warp-uniform)
  uint32_t  carry, cbits;   // trailing partial word, top aligned; its bits  (warp-uniform)
  uint32_t  flushed;        // words already written to global               (warp-uniform)
  uint8_t  *gptr;           // global address of word 0 of the stream

  uint32_t  cap;            // words the stage holds (plus 4 words of slack)
  uint32_t  limit;          // bytes there is room for at gptr; a flush past it is dropped and sets ovf
  uint32_t  ovf;            //                                                (warp-uniform)

  __device__ __forceinline__ void init(uint32_t *st, uint8_t *g, uint32_t capacity)
  { stage = st; nst = 0; carry = 0; cbits = 0; flushed = 0; gptr = g; cap = capacity;
    limit = 0xffffffffu; ovf = 0;
  }
  __device__ __forceinline__ uint32_t bitpos() const { return nst*32u + cbits; }       // in the stage
  __device__ __forceinline__ uint32_t total() const { return (flushed + nst)*32u + cbits; }

  // make room for `bits` more bits
  template <bool SWAP>
  __device__ __forceinline__ void reserve(uint32_t bits, int lane)
  { if (nst + ((cbits + bits + 31u) >> 5) + 1u > cap)
      { __syncwarp();
        if ((flushed + nst)*4u <= limit) copy_out<SWAP>(gptr + (size_t) flushed*4u,stage,nst*4u,lane);
        else                             ovf = 1;
        __syncwarp();
        flushed += nst; nst = 0;
      }
  }
};

// one lane's bit string inside a row: pieces are shifted through a 64-bit register; every word the
// lane completes is stored, what is left over joins its neighbours in finish()
struct LaneSink
{ uint64_t acc; uint32_t nacc, widx, fw;
  __device__ __forceinline__ void start(uint32_t pos)
  { widx = fw = pos >> 5; nacc = pos & 31u; acc = 0; }
  __device__ __forceinline__ void put(uint32_t *stage, uint32_t bits, uint32_t len)    // len <= 32
  { acc = (acc << len) | bits;
    nacc += len;
    if (nacc >= 32u)
      { nacc -= 32u;
        stage[widx++] = (uint32_t) (acc >> nacc);
      }
  }
  // warp-collective, for rounds in which EVERY lane contributed at least 32 bits: each lane then
  // completes the word it starts in, so the partial word at its end is completed by its upper
  // neighbour and by nobody else -- one shuffle instead of the segmented scan of finish()
  __device__ __forceinline__ void finish_wide(WarpBits &wb, int lane)
  { const uint32_t t = nacc ? (uint32_t) (acc << (32u - nacc)) : 0u;    // my trailing partial word
    uint32_t before = __shfl_up_sync(DX_FULL,t,1);                       // what is already in my first word
    if (lane == 0) before = wb.carry;
    if (before) wb.stage[fw] |= before;
    wb.carry = __shfl_sync(DX_FULL,t,31);
    wb.nst = __shfl_sync(DX_FULL,widx,31); wb.cbits = __shfl_sync(DX_FULL,nacc,31);
  }

  // warp-collective: merge the partial words, update the warp state
  __device__ __forceinline__ void finish(WarpBits &wb, int lane)
  { uint32_t t = nacc ? (uint32_t) (acc << (32u - nacc)) : 0u;          // my trailing partial word
    if (lane == 0 && widx == fw) t |= wb.carry;                          // still in the carry's word
    // segmented inclusive OR-scan keyed by the word the partial belongs to (keys ascend by lane)
    const uint32_t kprev = __shfl_up_sync(DX_FULL,widx,1);
    const uint32_t heads = __ballot_sync(DX_FULL,lane == 0 || kprev != widx);
    const int seg = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));   // first lane of my segment
    uint32_t sc = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
      { const uint32_t o = __shfl_up_sync(DX_FULL,sc,d);
        if (lane - d >= seg) sc |= o;
      }
    uint32_t before = __shfl_up_sync(DX_FULL,sc,1);                      // what is already in my first word
    if (lane == 0) before = wb.carry;
    if (widx != fw && before) wb.stage[fw] |= before;                    // I completed that word
    wb.carry = __shfl_sync(DX_FULL,sc,31);
    const uint32_t endw = __shfl_sync(DX_FULL,widx,31), endb = __shfl_sync(DX_FULL,nacc,31);
    wb.nst = endw; wb.cbits = endb;
    if (endb == 0) wb.carry = 0;
  }
};

