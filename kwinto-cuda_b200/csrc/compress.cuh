// compress.cuh -- device-side chain compression.
//
// The reference sorts option indices by (t, r, q, z, e, w) and walks the runs to build
// asset2pde / pde2asset (src/Pricer/kwFd1d.cpp:28-65): one PDE per chain, because s and k are not
// in the key.  On the host that sort costs more than the whole GPU march for large portfolios, so
// here the grouping is a hash join in HBM: three small kernels, no host round trip (the march
// kernel reads the PDE count from device memory).
//   1. chain_insert : open-addressing table keyed on the six fields; the first option to claim a
//      slot becomes the chain's representative, every option remembers its slot and bumps the
//      slot's member count;
//   2. chain_compact: every claimed slot takes the next PDE number and a segment of the member
//      list (two atomic counters);
//   3. chain_fill   : options drop their index into their chain's segment.
// PDE numbering and the order inside a segment depend on atomics, prices do not: a PDE's solution
// is a function of the key only and every option is interpolated independently.
#pragma once
#include "fd1d_common.cuh"

namespace kwfd1d {

constexpr uint32_t kEmptySlot = 0xffffffffu;

struct ChainTable {
    uint32_t* slot_rep;   // [cap] representative option of the chain in this slot, or kEmptySlot
    uint32_t* slot_cnt;   // [cap] members
    uint32_t* slot_pde;   // [cap] PDE number given by chain_compact
    uint32_t* opt_slot;   // [n]   slot of every option
    uint32_t* rep;        // [n]   out: pde -> representative option
    uint32_t* seg_start;  // [n]   out: pde -> first entry of its member list
    uint32_t* seg_cnt;    // [n]   out: pde -> members
    uint32_t* seg_fill;   // [n]   scratch cursors
    uint32_t* members;    // [n]   out: member lists
    uint32_t* counters;   // [0] = number of PDEs, [1] = member-list cursor
    uint32_t cap;         // power of two >= 2n
};

__device__ __forceinline__ uint64_t chain_mix(uint64_t h, uint64_t v)
{
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    return h;
}

// Bit pattern of a key field for hashing.  chain_same compares with ==, as the reference's std::tie comparison
// does (src/Pricer/kwFd1d.cpp:33-35), and -0.0 == +0.0: both zeros must hash alike or a chain with q = -0.0 and
// one with q = +0.0 would become two PDEs where the reference solves one.  x + 0.0 is +0.0 for either zero.
__device__ __forceinline__ uint64_t chain_bits(double x) { return (uint64_t)__double_as_longlong(x + 0.0); }

__device__ __forceinline__ uint32_t chain_hash(const kw_option& o)
{
    uint64_t h = 0x243f6a8885a308d3ull;
    h = chain_mix(h, chain_bits(o.t) * 0xff51afd7ed558ccdull);
    h = chain_mix(h, chain_bits(o.r) * 0xff51afd7ed558ccdull);
    h = chain_mix(h, chain_bits(o.q) * 0xff51afd7ed558ccdull);
    h = chain_mix(h, chain_bits(o.z) * 0xff51afd7ed558ccdull);
    h = chain_mix(h, ((uint64_t)o.e << 8) | (uint8_t)o.w);
    return (uint32_t)(h ^ (h >> 32));
}

// the reference's chain key (src/Pricer/kwFd1d.cpp:33-35)
__device__ __forceinline__ bool chain_same(const kw_option& l, const kw_option& r)
{
    return l.t == r.t && l.r == r.r && l.q == r.q && l.z == r.z && l.e == r.e && l.w == r.w;
}

__global__ void chain_reset_kernel(ChainTable T)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < T.cap) {
        T.slot_rep[i] = kEmptySlot;
        T.slot_cnt[i] = 0u;
    }
    if (i < 2) T.counters[i] = 0u;
}

__global__ void chain_insert_kernel(const kw_option* opts, uint32_t n, ChainTable T)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const kw_option o = load_option(opts + i);
    const uint32_t mask = T.cap - 1;
    uint32_t h = chain_hash(o) & mask;
    for (;;) {
        uint32_t cur = T.slot_rep[h];
        if (cur == kEmptySlot) cur = atomicCAS(&T.slot_rep[h], kEmptySlot, i);
        if (cur == kEmptySlot) break;  // claimed: option i represents a new chain
        // plain (not read-only-path) loads: the representative was written by the same launch's H2D only
        if (chain_same(o, load_option(opts + cur))) break;
        h = (h + 1) & mask;
    }
    T.opt_slot[i] = h;
    atomicAdd(&T.slot_cnt[h], 1u);
}

__global__ void chain_compact_kernel(ChainTable T)
{
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= T.cap) return;
    const uint32_t r = T.slot_rep[h];
    if (r == kEmptySlot) return;
    const uint32_t cnt = T.slot_cnt[h];
    const uint32_t p = atomicAdd(&T.counters[0], 1u);
    const uint32_t s = atomicAdd(&T.counters[1], cnt);
    T.slot_pde[h] = p;
    T.rep[p] = r;
    T.seg_start[p] = s;
    T.seg_cnt[p] = cnt;
    T.seg_fill[p] = 0u;
}

__global__ void chain_fill_kernel(uint32_t n, ChainTable T)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = T.slot_pde[T.opt_slot[i]];
    const uint32_t at = atomicAdd(&T.seg_fill[p], 1u);
    T.members[T.seg_start[p] + at] = i;
}

}  // namespace kwfd1d
