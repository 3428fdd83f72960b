// kernels.cuh — sm_100a kernels of the gxsv statevector engine.
//
// Every kernel is a streaming pass over (part of) the amplitude array: O(1)
// flops per 16-byte amplitude, so the bound is HBM bandwidth.
// The common shape is therefore: 256-thread CTAs, a persistent grid sized in
// multiples of the SM count, each thread issuing U independent 16-byte
// (double2) loads before the first use so that >=64 KB per SM are in flight,
// warp-coalesced 512-byte rows, block partials + a deterministic last-block
// reduction for the norms.
//
// Physical layout.  The live amplitudes occupy the sub-cube of the physical
// index space spanned by the *live* physical bits; dead bits ("holes", always
// at bit >= 8 so live runs are >= 4 KiB contiguous) are fixed to 0.  A kernel
// enumerates a dense rank index j and expands it to a physical index by
// inserting zero bits at a sorted list of positions (`Holes`): the holes of
// the layout plus the axis bits the kernel pairs over.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace gxk {

constexpr int TPB = 256;          // threads per CTA for every streaming kernel
constexpr int MAX_BITS = 48;

struct Holes {
  int n;
  unsigned char pos[MAX_BITS - 4];
};

struct Ctrl {            // device-resident control block
  double g;              // pending real normalisation: true psi = g * h(host) * stored
  unsigned int ticket;   // last-block election
  unsigned int pad;
};

constexpr int RES_VALUES = 16;
struct Res {             // one slot of the host-mapped result ring (256 B)
  double v[RES_VALUES];
  unsigned long long seq;
  unsigned long long pad[15];
};

__device__ __forceinline__ uint64_t expand(uint64_t x, const Holes& h) {
  for (int i = 0; i < h.n; ++i) {
    const uint64_t lo = x & ((1ull << h.pos[i]) - 1ull);
    x = ((x - lo) << 1) | lo;
  }
  return x;
}

__device__ __forceinline__ double2 cmul(const double2 a, const double2 b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// a*x + b*y
__host__ __device__ __forceinline__ double2 cmad2(const double2 a, const double2 x, const double2 b, const double2 y) {
  double re = a.x * x.x;
  re = fma(-a.y, x.y, re);
  re = fma(b.x, y.x, re);
  re = fma(-b.y, y.y, re);
  double im = a.x * x.y;
  im = fma(a.y, x.x, im);
  im = fma(b.x, y.y, im);
  im = fma(b.y, y.x, im);
  return make_double2(re, im);
}
__device__ __forceinline__ double cnorm2(const double2 a) { return fma(a.x, a.x, a.y * a.y); }

// streaming accesses: the state is touched once per pass and is far larger than L2
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 ld_amp(const double2* p) { return __ldcs(p); }
__device__ __forceinline__ void st_amp(double2* p, const double2 v) { __stcs(p, v); }

// ---------------------------------------------------------------------------
// deterministic grid reduction of NV doubles per thread
// ---------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /* NV*8 */) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) smem[k * 8 + w] = v[k];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double x = (l < TPB / 32) ? smem[k * 8 + l] : 0.0;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      v[k] = x;
    }
  }
}

enum { FIN_CONTRACT = 0, FIN_SCALED = 1, FIN_CONTRACT_DIST = 2, FIN_BATCH = 3, FIN_ONESHOT = 4 };

// Writes the block partial, elects the last block, which sums all partials in a
// fixed order and publishes the result to the host-mapped slot.
//   FIN_CONTRACT: v[0] = sum |stored'|^2 (the stored state now has that norm);
//                 ctrl->g = 1/sqrt(v[0]) so that g*stored is normalised.
//   FIN_SCALED  : v[k] *= post * g^2 (expectation values / norms of the true state).
//   FIN_CONTRACT_DIST: sharded state — publish the local sum only; g is set by the host after the
//                 all-reduce over ranks (gxsv_set_norm).
//   FIN_BATCH    : a batch of `post` projection stages: v[k] = norm^2 after stage k; g = 1/sqrt(v[post-1]).
//   FIN_ONESHOT  : same publication, but g <- 1 (sharded non-blocking mode: normalisation is set by the host).
template <int NV>
__device__ __forceinline__ void grid_finish(double (&acc)[NV], double* __restrict__ partials, Ctrl* ctrl,
                                            Res* res, unsigned long long seq, int mode, double post, double g,
                                            const double* nscale = nullptr) {
  __shared__ double red[NV * 8];
  __shared__ bool is_last;
  block_sum<NV>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) partials[(size_t)blockIdx.x * NV + k] = acc[k];
    __threadfence();
    const unsigned int t = atomicAdd(&ctrl->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double tot[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) tot[k] = 0.0;
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += TPB) {
#pragma unroll
    for (int k = 0; k < NV; ++k) tot[k] += __ldcg(&partials[(size_t)b * NV + k]);
  }
  block_sum<NV>(tot, red);
  if (threadIdx.x == 0) {
    ctrl->ticket = 0;
    if (mode == FIN_CONTRACT) {
      ctrl->g = tot[0] > 0.0 ? 1.0 / sqrt(tot[0]) : 1.0;
      res->v[0] = tot[0] * (nscale ? nscale[0] : 1.0);
    } else if (mode == FIN_CONTRACT_DIST) {
      res->v[0] = tot[0];
    } else if (mode == FIN_BATCH) {
      // stages may keep a scalar f out of the stored amplitudes (Stage kinds 0 / 1): the published norms are those of
      // the unfactored state, g normalises what is stored
#pragma unroll
      for (int k = 0; k < NV; ++k) res->v[k] = tot[k] * (nscale ? nscale[k] : 1.0);
      // a zero norm (an impossible branch was taken) keeps g finite: the stored zeros stay zeros instead of NaNs, and the
      // deferred check raises the reference's RuntimeError as the first error the caller sees
      ctrl->g = tot[(int)post - 1] > 0.0 ? 1.0 / sqrt(tot[(int)post - 1]) : 1.0;
    } else if (mode == FIN_ONESHOT) {
      // sharded, non-blocking: publish the local norms; the pending factor g has been consumed by this pass and
      // stays 1 until the host sets it again from the all-reduced norm (gxsv_set_norm)
#pragma unroll
      for (int k = 0; k < NV; ++k) res->v[k] = tot[k] * (nscale ? nscale[k] : 1.0);
      ctrl->g = 1.0;
    } else {
#pragma unroll
      for (int k = 0; k < NV; ++k) res->v[k] = tot[k] * post * g * g;
    }
    __threadfence_system();
    *((volatile unsigned long long*)&res->seq) = seq;
  }
}

// ---------------------------------------------------------------------------
// N — append one qubit on a free physical bit `hbit`.
//   MODE 0: psi[i|h] = psi[i]            (|+>: the 1/sqrt2 is deferred to the host scalar)
//   MODE 1: psi[i|h] = f1*psi[i]         (s0 deferred to the host scalar, f1 = s1/s0)
//   MODE 2: psi[i|h] = f1*psi[i]; psi[i] = f0*psi[i]
// Reads S, writes S (MODE 2: 2S); in place and hazard free because thread i is
// the only one touching i and i|h.
// ---------------------------------------------------------------------------
template <int U, int MODE>
__global__ void __launch_bounds__(TPB) k_fill(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                              const int hbit, const double2 f0, const double2 f1,
                                              const uint64_t nlive) {
  const uint64_t chunk = (uint64_t)TPB * U;
  const uint64_t hoff = 1ull << hbit;
  for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < nlive; base += (uint64_t)gridDim.x * chunk) {
    double2 a[U];
    uint64_t s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      s[u] = expand(j, hs);
      if (j < nlive) a[u] = ld_amp(psi + s[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      if (j < nlive) {
        if (MODE == 0) {
          st_amp(psi + s[u] + hoff, a[u]);
        } else {
          st_amp(psi + s[u] + hoff, cmul(f1, a[u]));
          if (MODE == 2) st_amp(psi + s[u], cmul(f0, a[u]));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// E-run — one star of CZs sharing the pivot qubit m: negate psi[x] where
// x_m = 1 and parity(x & nbrs) = 1.  Only that quarter of the amplitudes is
// enumerated: j runs over the other n-2 bits, bit m is forced to 1 and the
// highest neighbour bit h is solved from the parity of the remaining ones.
// Reads S/4, writes S/4.
// ---------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(TPB) k_cz_star(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                                 const int mbit, const int hbit, const uint64_t nbr_rest,
                                                 const uint64_t nquarter) {
  const uint64_t chunk = (uint64_t)TPB * U;
  for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < nquarter; base += (uint64_t)gridDim.x * chunk) {
    double2 a[U];
    uint64_t s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      uint64_t x = expand(j, hs) | (1ull << mbit);
      const uint64_t par = (uint64_t)(__popcll(x & nbr_rest) & 1);
      x |= (par ^ 1ull) << hbit;
      s[u] = x;
      if (j < nquarter) a[u] = ld_amp(psi + x);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      if (j < nquarter) st_amp(psi + s[u], make_double2(-a[u].x, -a[u].y));
    }
  }
}

// General diagonal pass for E-runs that are not a single star:
// sign(x) = (-1)^{sum_e x_a x_b}  (+ linear terms in zmask).  Full read, writes only flipped amplitudes.
struct EdgeList {
  int n;
  unsigned char a[64];
  unsigned char b[64];
};
template <int U>
__global__ void __launch_bounds__(TPB) k_diag_edges(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                                    const __grid_constant__ EdgeList el, const uint64_t zmask,
                                                    const uint64_t nlive) {
  const uint64_t chunk = (uint64_t)TPB * U;
  for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < nlive; base += (uint64_t)gridDim.x * chunk) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      if (j < nlive) {
        const uint64_t x = expand(j, hs);
        unsigned int par = (unsigned int)__popcll(x & zmask);
        for (int e = 0; e < el.n; ++e) par += (unsigned int)((x >> el.a[e]) & (x >> el.b[e]) & 1ull);
        if (par & 1u) {
          const double2 a = ld_amp(psi + x);
          st_amp(psi + x, make_double2(-a.x, -a.y));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// M — project + compact + reduce in one pass:
//   out[j] = g*(c0*psi[s0(j)] + c1*psi[s0(j)|p]),  norm2 = sum |out|^2
// s0 inserts the layout holes and the measured bit p (hs); the result is
// written at d(j) (hd).  `mq` (physical-bit mask) carries CZs still queued
// between the measured qubit and other qubits: they are applied on the fly.  TILE = false: d == s0, the measured bit becomes a
// hole (p >= 8).  TILE = true: p is a low bit; each CTA owns one aligned tile
// of the lowest K live bits, reads all 2^K inputs into registers, barriers,
// and writes the 2^(K-1) outputs into the lower half of its own tile, so the
// hole moves up to the tile's top bit and low bits never fragment.
// Reads S, writes S/2.
// ---------------------------------------------------------------------------
template <int U, bool TILE>
__global__ void __launch_bounds__(TPB) k_contract(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                                  const __grid_constant__ Holes hd, const int pbit, double2 c0,
                                                  double2 c1, const uint64_t mq, const uint64_t nout,
                                                  double* __restrict__ partials, Ctrl* ctrl, Res* res,
                                                  const unsigned long long seq, const int fin_mode) {
  const double g = ctrl->g;
  c0.x *= g; c0.y *= g; c1.x *= g; c1.y *= g;
  const uint64_t chunk = (uint64_t)TPB * U;
  const uint64_t poff = 1ull << pbit;
  double acc[1] = {0.0};
  for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < nout; base += (uint64_t)gridDim.x * chunk) {
    double2 a[U], b[U];
    uint64_t s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      s[u] = expand(j, hs);
      if (j < nout) {
        a[u] = ld_amp(psi + s[u]);
        b[u] = ld_amp(psi + s[u] + poff);
      }
    }
    if (TILE) __syncthreads();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      if (j < nout) {
        // queued CZs between the measured qubit and the qubits in mq: sign on the x_p = 1 amplitude
        const double sg = (__popcll(s[u] & mq) & 1) ? -1.0 : 1.0;
        const double2 r = cmad2(c0, a[u], c1, make_double2(sg * b[u].x, sg * b[u].y));
        acc[0] += cnorm2(r);
        const uint64_t d = TILE ? expand(j, hd) : s[u];
        st_amp(psi + d, r);
      }
    }
  }
  grid_finish<1>(acc, partials, ctrl, res, seq, fin_mode, 1.0, g);
}

// ---------------------------------------------------------------------------
// Lazy N,E..,M fusion.  A qubit v that was prepared (state (s0,s1)) and entangled
// with the measured qubit q but never materialised takes q's physical bit:
//   t0 = psi[x, q=0],  t1 = sq(x) * psi[x, q=1]            sq = (-1)^{parity(x & mq)}
//   psi'[x, v=0] =          m00*t0 + m01*t1                m = [[s0 c0, s0 c1], [s1 c0, -s1 c1]]
//   psi'[x, v=1] = sv(x) * (m10*t0 + m11*t1)               sv = (-1)^{parity(x & mv)}
// with <c| the measurement bra, mq / mv the queued CZ partners of q / v.
// One in-place butterfly: reads S, writes S (instead of N 2S + E S/2 + M 3S at the
// wider width), no hole is created and the norm is reduced in the same pass.
// ---------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(TPB) k_fused(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                               const int pbit, double2 m00, double2 m01, double2 m10, double2 m11,
                                               const uint64_t mq, const uint64_t mv, const uint64_t npair,
                                               double* __restrict__ partials, Ctrl* ctrl, Res* res,
                                               const unsigned long long seq, const int fin_mode,
                                               const double nscale0) {
  const double g = ctrl->g;
  m00.x *= g; m00.y *= g; m01.x *= g; m01.y *= g; m10.x *= g; m10.y *= g; m11.x *= g; m11.y *= g;
  const uint64_t chunk = (uint64_t)TPB * U;
  const uint64_t poff = 1ull << pbit;
  double acc[1] = {0.0};
  for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < npair; base += (uint64_t)gridDim.x * chunk) {
    double2 a[U], b[U];
    uint64_t s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      s[u] = expand(j, hs);
      if (j < npair) {
        a[u] = ld_amp(psi + s[u]);
        b[u] = ld_amp(psi + s[u] + poff);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      if (j < npair) {
        const double sq = (__popcll(s[u] & mq) & 1) ? -1.0 : 1.0;
        const double sv = (__popcll(s[u] & mv) & 1) ? -1.0 : 1.0;
        const double2 t1 = make_double2(sq * b[u].x, sq * b[u].y);
        const double2 r0 = cmad2(m00, a[u], m01, t1);
        double2 r1 = cmad2(m10, a[u], m11, t1);
        r1.x *= sv; r1.y *= sv;
        acc[0] += cnorm2(r0) + cnorm2(r1);
        st_amp(psi + s[u], r0);
        st_amp(psi + s[u] + poff, r1);
      }
    }
  }
  const double ns[1] = {nscale0};   // the matrix may have a scalar taken out (a batch of one, see Stage): published norm *= |f|^2
  grid_finish<1>(acc, partials, ctrl, res, seq, fin_mode, 1.0, g, ns);
}

// ---------------------------------------------------------------------------
// Batched lazy projections (non-strict mode).  Up to MAX_STAGES consecutive fused N,E..,M butterflies are
// queued on the host and executed in ONE pass: the stages act on at most A distinct physical bits, a thread
// owns the 2^A amplitudes of one group in registers and applies the stages in order (each stage is the
// k_fused butterfly on its own axis, sign masks evaluated on the current index bits), accumulating the
// norm^2 after every stage.  K stages cost 2 S instead of 2 K S.  Measured alternatives that did not help: 4 axes per
// pass (227 registers, 1 CTA/SM), a software-pipelined loop prefetching the next group (122 registers, 2 CTAs/SM) and
// a cheaper butterfly for |+> partners (u +- w; the extra branch costs more than the 8 multiplies it saves), and
// hoisting the sign parities out of the pair loop with sign-bit XORs instead of selects (same rate, 32 B of spills:
// the pass is not issue bound although ncu shows the issue slots 65 % busy), and staging the NEXT group of every
// thread in its own shared-memory column with 16-byte cp.async.cg copies while the current group is computed (bytes in
// flight during the FP64 phase, no extra registers: 0.418 ms per pass instead of 0.371 for 3 axes, 0.516 instead of
// 0.473 for 4 axes — the LDGSTS path streams worse than plain 16-byte evict-first loads), and a 4-axis pass that keeps
// half of every group in the thread's shared-memory column to stay at 3 CTAs per SM (80 registers, 104 B of spills:
// 0.540 ms per pass);
// with 3 stages the pass runs at ~0.84 of the copy rate with the FP64 pipe ~45 % busy under the 1 kW power cap.
// ---------------------------------------------------------------------------
constexpr int MAX_STAGES = 12;
// One fused (N, E.., M) stage on one axis.  With <c| the measurement bra (host scalar folded in), (s0, s1) the state of
// the partner v and sq / sv the CZ signs of the measured qubit / of v (k_fused above):
//   r0 = s0 (c0 t0 + c1 t1),   r1 = sv s1 (c0 t0 - c1 t1),      t0 = psi[.., 0], t1 = sq psi[.., 1].
// kind 2 (general): the four products are in m00..m11 and the stage costs two complex 2-term dot products.
// kind 0 / 1 (s1 = +-s0, every |+> / |-> partner): the scalar f = s0 c0 (kind 0, |c0| >= |c1|) or f = s0 c1 (kind 1) is
//   taken out of the stage — the host keeps its phase, the published norms are multiplied by |f|^2 (Batch::nscale) — so that
//   r0 = t0 + a t1, r1 = +-sv (t0 - a t1), a = c1/c0      (kind 0)       six fused multiply-adds per pair
//   r0 = a t0 + t1, r1 = +-sv (a t0 - t1), a = c0/c1      (kind 1)       (stage_pairs; the general form costs 21 operations).
// The CZ signs are sign-bit XORs: the parity of the thread's fixed index bits is taken once per stage and combined with
// a host-made table (qtab / vtab: parity contributed by the register-axis bits of the k-th pair, one byte per pair,
// flag in the byte's top bit so that one byte permute yields the 0 / 0x80000000 mask of a pair).
struct Stage {
  double2 m00, m01, m10, m11;    // the unfactored 2x2 (kind 2; also what a batch of one hands to k_fused)
  double2 a;                     // kinds 0 / 1: the one multiplier left after taking f out
  unsigned long long mq, mv;
  int axis;                      // batch axis index
  unsigned char kind;            // 0 / 1 / 2, see above
  unsigned char neg;             // kinds 0/1: s1 = -s0
  unsigned char acc;             // 1 = accumulate the norm^2 after this stage (0: the host knows it analytically)
  unsigned char pad;
  unsigned long long qtab, vtab; // set when the pass is launched (depend on the register assignment of the axes):
                                 // byte k = 0x80 iff the k-th amplitude pair of the stage's axis takes the sign
};
struct Batch {
  int nstages;
  int naxes;
  unsigned char axbit[8];
  double nscale[MAX_STAGES];     // prod_{i <= k} |f_i|^2: stored norm -> norm of the unfactored state after stage k
  Stage st[MAX_STAGES];
};

// (the stage arithmetic below is __host__ __device__ so that gxsv_debug_stage can run it on the CPU for tests/test_stage_math.py;
// the device code is unchanged by that)
__host__ __device__ __forceinline__ double flip_sign(const double x, const unsigned int s31) {   // s31 = 0 or 0x80000000
#ifdef __CUDA_ARCH__
  return __hiloint2double(__double2hiint(x) ^ (int)s31, __double2loint(x));
#else
  unsigned long long u;
  memcpy(&u, &x, sizeof u);
  u ^= (unsigned long long)s31 << 32;
  double r;
  memcpy(&r, &u, sizeof r);
  return r;
#endif
}
// one byte of a sign table -> 0 / 0x80000000 (device: a single byte permute; sel = 0x0444 | (k << 12) moves byte k of t to
// the top byte and clears the rest)
__host__ __device__ __forceinline__ unsigned int sign_of_byte(const unsigned int t, const int k) {
#ifdef __CUDA_ARCH__
  return __byte_perm(t, 0u, 0x0444 | (k << 12));
#else
  return ((t >> (8 * k)) & 0xffu) << 24;
#endif
}

// The 2^(NB-1) amplitude pairs of one stage on register axis AX.  KIND as in Stage; ACC = accumulate the norm^2 of
// the results (two chains); HASV = the new qubit's half takes a sign (CZs of the partner with materialised qubits, or a
// |-> partner).  For kinds 0 / 1 with |a| = 1 — every XY-plane measurement of a qubit whose partner was prepared in
// |+> — the norm after the stage is exactly 2 |in|^2 whatever the data (the outcome probability of such a measurement
// is 1/2), so the host asks for the accumulation only where it cannot know the answer (Stage::acc).
// FP64 work of a kind 0 / 1 pair: SIX fused multiply-adds — r1 = t0 - a t1 as two chains of two, then the other output
// from the first, r0 = 2 t0 - r1 (kind 1: r0 = a t0 + t1, r1 = r0 - 2 t1) — against 2 DMUL + 2 DFMA + 4 DADD for the
// textbook form; the CZ sign of the measured qubit is XORed into t1 in place (two LOP3, no register moves), and the
// batched path queues no partner signs at all (gxsv_project_qubit leaves those CZs queued until one of their ends is
// measured, where they are the free sq of that stage), so the common stage is 6 DFMA + 2 LOP3 + 1 PRMT per pair.
// UNIQ (kind 0 only): no CZ partner of the measured qubit sits on another register axis, so the sign of t1 is the same
// for all pairs of the thread and is folded into the multiplier once (no per-pair sign work at all: 6 DFMA per pair).
template <int NB, int AX, int KIND, bool ACC, bool HASV, bool UNIQ = false>
__host__ __device__ __forceinline__ void stage_pairs(double2 (&v)[1 << NB], const Stage& st, double2 a, const unsigned long long tq64,
                                            const unsigned long long tv64, double& nrm0, double& nrm1) {
  if (UNIQ) {
    const unsigned int s = sign_of_byte((unsigned int)tq64, 0);
    a.x = flip_sign(a.x, s);
    a.y = flip_sign(a.y, s);
  }
#pragma unroll
  for (int lo = 0; lo < (1 << NB); ++lo) {
    if ((lo >> AX) & 1) continue;
    const int hi = lo | (1 << AX);
    const int k = ((lo >> (AX + 1)) << AX) | (lo & ((1 << AX) - 1));   // index of this pair (folded at compile time)
    const unsigned int tq = (k < 4) ? (unsigned int)tq64 : (unsigned int)(tq64 >> 32);
    const unsigned int sq = UNIQ ? 0u : sign_of_byte(tq, k & 3);
    unsigned int sv = 0u;
    if (HASV) {
      const unsigned int tv = (k < 4) ? (unsigned int)tv64 : (unsigned int)(tv64 >> 32);
      sv = sign_of_byte(tv, k & 3);
    }
    const double t1x = UNIQ ? v[hi].x : flip_sign(v[hi].x, sq), t1y = UNIQ ? v[hi].y : flip_sign(v[hi].y, sq);
    // order chosen so that every result can take the registers of the operand it replaces (two temporaries per pair):
    // the half that needs both components of one input first, into that input's slot, then the other half in place
    if (KIND == 0) {
      const double ux = fma(a.y, t1y, v[lo].x), uy = fma(-a.y, t1x, v[lo].y);
      v[hi].x = fma(-a.x, t1x, ux);                 // r1 = t0 - a t1
      v[hi].y = fma(-a.x, t1y, uy);
      v[lo].x = fma(2.0, v[lo].x, -v[hi].x);        // r0 = t0 + a t1 = 2 t0 - r1
      v[lo].y = fma(2.0, v[lo].y, -v[hi].y);
      if (HASV) { v[hi].x = flip_sign(v[hi].x, sv); v[hi].y = flip_sign(v[hi].y, sv); }
    } else if (KIND == 1) {
      const double ux = fma(-a.y, v[lo].y, t1x), uy = fma(a.y, v[lo].x, t1y);
      v[lo].x = fma(a.x, v[lo].x, ux);              // r0 = a t0 + t1
      v[lo].y = fma(a.x, v[lo].y, uy);
      v[hi].x = fma(-2.0, t1x, v[lo].x);            // r1 = a t0 - t1 = r0 - 2 t1
      v[hi].y = fma(-2.0, t1y, v[lo].y);
      if (HASV) { v[hi].x = flip_sign(v[hi].x, sv); v[hi].y = flip_sign(v[hi].y, sv); }
    } else {
      const double2 t1 = make_double2(t1x, t1y);
      const double2 r0 = cmad2(st.m00, v[lo], st.m01, t1);
      double2 r1 = cmad2(st.m10, v[lo], st.m11, t1);
      v[hi].x = flip_sign(r1.x, sv);
      v[hi].y = flip_sign(r1.y, sv);
      v[lo] = r0;
    }
    if (ACC) {
      nrm0 = fma(v[lo].x, v[lo].x, nrm0); nrm1 = fma(v[lo].y, v[lo].y, nrm1);
      nrm0 = fma(v[hi].x, v[hi].x, nrm0); nrm1 = fma(v[hi].y, v[hi].y, nrm1);
    }
  }
}

// Apply stage `st` on register axis AX of the 2^NB amplitudes a thread holds; gfixed = the thread's fixed physical index
// bits (their parity in st.mq / st.mv decides the thread-wide part of the signs).  Adds the norm^2 of the 2^NB results to nrm0 / nrm1 (the caller keeps
// the sum only when st.acc is set; the rare variants always compute it — fewer code paths).
template <int NB, int AX>
__device__ __forceinline__ void stage_on_axis(double2 (&v)[1 << NB], const Stage& st, const uint64_t gfixed,
                                              double& nrm0, double& nrm1) {
  // byte k of tq / tv = 0x80 iff the k-th amplitude pair takes the sign
  const unsigned int gpar_q = __popcll(gfixed & st.mq);
  const unsigned long long tq = st.qtab ^ ((gpar_q & 1u) ? 0x8080808080808080ull : 0ull);
  if (st.kind == 0 && st.mv == 0ull && !st.neg) {          // every stage of a transpiled pattern
    if (st.acc) stage_pairs<NB, AX, 0, true, false>(v, st, st.a, tq, 0ull, nrm0, nrm1);
    else if (st.qtab == 0ull) stage_pairs<NB, AX, 0, false, false, true>(v, st, st.a, tq, 0ull, nrm0, nrm1);
    else stage_pairs<NB, AX, 0, false, false>(v, st, st.a, tq, 0ull, nrm0, nrm1);
    return;
  }
  const unsigned int gpar_v = __popcll(gfixed & st.mv);
  const unsigned long long tv = st.vtab ^ (((gpar_v ^ st.neg) & 1u) ? 0x8080808080808080ull : 0ull);
  if (st.kind == 0) stage_pairs<NB, AX, 0, true, true>(v, st, st.a, tq, tv, nrm0, nrm1);
  else if (st.kind == 1) stage_pairs<NB, AX, 1, true, true>(v, st, st.a, tq, tv, nrm0, nrm1);
  else stage_pairs<NB, AX, 2, true, true>(v, st, st.a, tq, tv, nrm0, nrm1);
}

// Batch on at most three physical bits: a thread owns the 2^A amplitudes of one group in registers and applies the
// stages in order (64 registers, 4 CTAs per SM; the per-thread stage accumulators live in shared memory).
// CTAs take chunks of 2^glog consecutive 256-group blocks (glog = 0: plain grid stride).
template <int A>
__global__ void __launch_bounds__(TPB, 4) k_fused_batch(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                                       const __grid_constant__ Batch bt, const uint64_t ngroups,
                                                       const int glog, double* __restrict__ partials, Ctrl* ctrl,
                                                       Res* res, const unsigned long long seq, const int fin_mode) {
  constexpr int D = 1 << A;
  const double g = ctrl->g;
  // offset of register slot b = OR of the axis bits set in b; only the A single-bit offsets are kept live
  uint64_t ax[A];
#pragma unroll
  for (int a = 0; a < A; ++a) ax[a] = 1ull << bt.axbit[a];
  auto off_of = [&](int b) {
    uint64_t o = 0;
#pragma unroll
    for (int a = 0; a < A; ++a)
      if ((b >> a) & 1) o |= ax[a];
    return o;
  };
  __shared__ double sacc[MAX_STAGES][TPB];
#pragma unroll
  for (int k = 0; k < MAX_STAGES; ++k) sacc[k][threadIdx.x] = 0.0;
  const uint64_t nblk = (ngroups + TPB - 1) / TPB;
  for (uint64_t blk0 = (uint64_t)blockIdx.x << glog; blk0 < nblk; blk0 += (uint64_t)gridDim.x << glog)
  for (uint64_t blk = blk0; blk < blk0 + (1ull << glog); ++blk) {
    const uint64_t j = blk * TPB + threadIdx.x;
    if (j >= ngroups) break;
    const uint64_t base = expand(j, hs);
    double2 v[D];
#pragma unroll
    for (int b = 0; b < D; ++b) {
      const double2 a = ld_amp(psi + base + off_of(b));
      v[b] = make_double2(g * a.x, g * a.y);       // the pending normalisation, once per pass
    }
    for (int s = 0; s < bt.nstages; ++s) {
      const Stage& st = bt.st[s];
      double n0 = 0.0, n1 = 0.0;
      const int la = st.axis;
      if (la == 0) stage_on_axis<A, 0>(v, st, base, n0, n1);
      if (A > 1 && la == 1) stage_on_axis<A, (A > 1 ? 1 : 0)>(v, st, base, n0, n1);
      if (A > 2 && la == 2) stage_on_axis<A, (A > 2 ? 2 : 0)>(v, st, base, n0, n1);
      if (st.acc) sacc[s][threadIdx.x] += n0 + n1;
    }
#pragma unroll
    for (int b = 0; b < D; ++b) st_amp(psi + base + off_of(b), v[b]);
  }
  double acc[MAX_STAGES];
#pragma unroll
  for (int k = 0; k < MAX_STAGES; ++k) acc[k] = sacc[k][threadIdx.x];
  grid_finish<MAX_STAGES>(acc, partials, ctrl, res, seq, fin_mode, (double)bt.nstages, g, bt.nscale);
}

// ---------------------------------------------------------------------------
// Batched lazy projections on MORE than three physical bits (4 <= A <= 7): the register file cannot hold 2^A amplitudes
// per thread, so a batch whose stages touch up to 7 distinct bits is executed in PHASES on TILES.
//   tile   = 2^WT amplitudes owned by a GROUP of TW warps (TW = 1: WT = 5 + R, a private warp tile; TW = 4: WT = 7 + R,
//            two groups per CTA): the 2^A combinations of the batch axes x the 2^F combinations of the F = WT - A lowest
//            live physical bits that are not axes.  The tile-index bits are thus WT physical bit positions.
//   layout = which R of those bits a thread keeps in registers (2^R amplitudes) and which thread-id bit stands for each
//            of the others.  Everything is tabulated by the host per layout (TileBatch::gbits / gtid for global
//            offsets, sbits / stid for slots of the group's shared-memory tile): the kernel ORs / XORs table entries.
//   phase  = a layout + the consecutive stages whose axes are register bits in it.  Between two layouts the group
//            transposes through its shared tile (only the group synchronises: __syncwarp, or a named barrier over
//            32 TW threads — groups and CTAs stay independent, so the FP64 work of one overlaps the loads of others).
//   I/O    = the first layout is the one the tile is LOADED in, the last one the one it is STORED in.  When the three
//            lowest tile bits are all batch axes (QFT patterns: 4 - 6 of the axes of a pass on physical bits 0 - 5)
//            the lanes of a direct layout would be >= 128 bytes apart; the host then puts an extra stage-less layout in
//            front and at the end whose thread bits are the LOWEST bits of the tile — axes or not — so that every
//            warp-wide access is one contiguous 512-byte run; otherwise the lanes run over the dense bits of the first /
//            last phase directly (two more transposes cost 8 % on config 2).  Measured before this (config 4, 32 live
//            qubits): such passes at 0.42 - 0.61 of the copy rate against 0.90 for passes on high bits.
// HBM traffic is one read and one write of the register per pass whatever the number of stages and layouts.
// Measured on the way (profiles/r02_summary.md): one 2^11 tile per CTA with __syncthreads between phases (1.37 ms for 8
// stages on 5 axes at 27 qubits), private warp tiles only (runs of 64 bytes on 7 axes: 0.58 of the copy rate on
// config 3 against 0.80 with four warps per tile), tiles shared by all 8 warps (9 % slower than 4), an L2 prefetch of
// the next tile (prefetch.global.L2: within +-2 %), streaming (ld.global.cs) loads (1 - 2 % slower than L1-allocating).
// ---------------------------------------------------------------------------
constexpr int TILE_LOG2 = 11;      // amplitudes per CTA iteration of k_fused_tma
constexpr int MAX_TILE_AXES = 7;
constexpr int MAX_PHASES = 4;      // phases with stages
constexpr int MAX_LAYOUTS = MAX_PHASES + 2;
constexpr int MAX_THREAD_BITS = 8; // log2 TPB
enum : unsigned int { SV_UNIQ = 0, SV_FAST = 1, SV_FAST_ACC = 2, SV_K0_SIGNED = 3, SV_K1 = 4, SV_GENERAL = 5 };
struct StageHot {
  double2 a;                    // Stage::a
  unsigned long long qtab, mq;  // Stage::qtab, Stage::mq
};
struct TileBatch {
  int nstages, naxes, nlayouts, F;
  unsigned char lax[MAX_STAGES];              // register slot (0..R-1) of the axis of every stage inside its phase
  unsigned char s_end[MAX_LAYOUTS + 2];       // stages [s_end[p-1], s_end[p]) run in layout p
  unsigned char code[MAX_STAGES + 4];         // 32 * lax + 16 * (accumulate the norm) + variant (SV_*), see stage_by_variant
  StageHot hot[MAX_STAGES + 1];               // the operands of the common variants, packed (one spare entry: fetched ahead)
  unsigned long long gload[16], gstore[16];              // BYTE offsets of register slot b in the first / last layout
  unsigned long long gbits[MAX_LAYOUTS][16];             // physical offset of register slot b in layout p
  unsigned long long gtid[MAX_LAYOUTS][MAX_THREAD_BITS]; // physical offset contributed by bit i of the thread's index in its group
  unsigned int sbits[MAX_LAYOUTS][16];                   // the same two as (swizzled) BYTE offsets inside the group's shared tile
  unsigned int stid[MAX_LAYOUTS][MAX_THREAD_BITS];
  double nscale[MAX_STAGES];
  Stage st[MAX_STAGES];
};

// One host-made code per stage (TileBatch::code: register axis, variant and the accumulate flag in one byte) and the
// three operands every common variant needs (multiplier, sign table, sign mask) in a packed table that the kernel
// fetches one stage ahead.  Without it a stage began with a chain of five dependent constant loads (register axis ->
// kind -> flags -> table -> multiplier), each an indexed LDC round trip with only four warps per scheduler to hide it
// (ncu, config 3: short-scoreboard stalls 1.9 - 3.3 per issued instruction).
template <int NB, int AX>
__device__ __forceinline__ void stage_by_variant(double2 (&v)[1 << NB], const unsigned int variant, const Stage& st,
                                                 const double2 a, const unsigned long long tq, const uint64_t gfixed,
                                                 double& nrm0, double& nrm1) {
  if (variant == SV_UNIQ) return stage_pairs<NB, AX, 0, false, false, true>(v, st, a, tq, 0ull, nrm0, nrm1);
  if (variant == SV_FAST) return stage_pairs<NB, AX, 0, false, false>(v, st, a, tq, 0ull, nrm0, nrm1);
  if (variant == SV_FAST_ACC) return stage_pairs<NB, AX, 0, true, false>(v, st, a, tq, 0ull, nrm0, nrm1);
  const unsigned int gpar_v = __popcll(gfixed & st.mv);
  const unsigned long long tv = st.vtab ^ (((gpar_v ^ st.neg) & 1u) ? 0x8080808080808080ull : 0ull);
  if (variant == SV_K0_SIGNED) stage_pairs<NB, AX, 0, true, true>(v, st, a, tq, tv, nrm0, nrm1);
  else if (variant == SV_K1) stage_pairs<NB, AX, 1, true, true>(v, st, a, tq, tv, nrm0, nrm1);
  else stage_pairs<NB, AX, 2, true, true>(v, st, a, tq, tv, nrm0, nrm1);
}

// R = register bits per layout: 3 (2^3 amplitudes per thread, 64 registers, 4 CTAs per SM) or 4 (2^4 amplitudes per
// thread, 128 registers, 2 CTAs per SM: fewer phases).  TW = warps per group (1, 4 or 8).  SP = the stage operands are
// staged in shared memory and fetched one stage ahead by asm-volatile loads (which neither the compiler nor ptxas sinks
// to the first use); SP = false reads them from the kernel parameters, where the fetch of stage s + 1 ends up at the
// bottom of stage s whatever the source order.
template <int R, int TW, bool SP>
__global__ void __launch_bounds__(TPB, (R == 3) ? 4 : 2) k_fused_tile(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                                      const __grid_constant__ TileBatch bt, const uint64_t ntiles,
                                                      const int glog, double* __restrict__ partials, Ctrl* ctrl,
                                                      Res* res, const unsigned long long seq, const int fin_mode) {
  constexpr int LTW = (TW == 1) ? 0 : (TW == 2) ? 1 : (TW == 4) ? 2 : 3;
  constexpr int WT = 5 + R + LTW;                   // log2 amplitudes per group tile
  constexpr int NTB = 5 + LTW;                      // thread-index bits of a group
  constexpr int GT = 32 * TW;                       // threads per group
  constexpr int NG = TPB / GT;                      // groups per CTA
  static_assert(NG == 1 || NG == 2 || TW == 1, "named barriers are written out for two groups");
  constexpr int D = 1 << R;
  extern __shared__ double2 tile_all[];             // NG groups x 2^WT amplitudes (dynamic) = 2^(8+R) amplitudes per CTA
  __shared__ double sacc[MAX_STAGES][TPB];
  __shared__ __align__(16) StageHot shot[MAX_STAGES + 1];
  __shared__ unsigned int scode[MAX_STAGES + 4];
  const int F = bt.F;
  const int tid = threadIdx.x, tg = tid & (GT - 1), grp = tid / GT;
  char* tile = reinterpret_cast<char*>(tile_all + ((size_t)grp << WT));   // addressed by XORs of tabulated byte offsets
  auto group_sync = [&]() {
    if (TW == 1) __syncwarp();
    else if (NG == 1) __syncthreads();
    else if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(GT) : "memory");   // immediate ids: a register id makes
    else asm volatile("bar.sync 2, %0;" ::"n"(GT) : "memory");                  // ptxas reserve all 16 barriers
  };
  const double g = ctrl->g;
  const int nlayouts = bt.nlayouts;
#pragma unroll
  for (int k = 0; k < MAX_STAGES; ++k) sacc[k][tid] = 0.0;
  if (SP) {
    if (tid <= MAX_STAGES) { shot[tid] = bt.hot[tid]; scode[tid] = bt.code[tid]; }
    __syncthreads();
  }
  // operands of stage k (SP: from shared memory, pinned where it is written)
  auto fetch = [&](int k, unsigned int& c, double2& am, unsigned long long& qt, unsigned long long& m) {
    if (SP) {
      const unsigned int ha = smem_u32(&shot[k]), ca = smem_u32(&scode[k]);
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(am.x), "=d"(am.y) : "r"(ha));
      asm volatile("ld.shared.v2.u64 {%0, %1}, [%2+16];" : "=l"(qt), "=l"(m) : "r"(ha));
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c) : "r"(ca));
    } else {
      c = bt.code[k];
      am = bt.hot[k].a;
      qt = bt.hot[k].qtab;
      m = bt.hot[k].mq;
    }
  };
  // this thread's part of the addresses in layout p: uniform table entries selected by the bits of its index in the group
  // (a table indexed by the thread itself — a per-thread index into the kernel parameters — faulted on sm_100a with
  // CUDA 12.9 although the emulated addresses were all in range; the uniform form costs a few predicated ORs)
  auto thread_off = [&](int p, uint64_t& go, unsigned int& so) {
    go = 0;
    so = 0;
#pragma unroll
    for (int i = 0; i < NTB; ++i)
      if ((tg >> i) & 1) { go |= bt.gtid[p][i]; so ^= bt.stid[p][i]; }
  };
  // the i-th CTA tile of this CTA: chunks of 2^glog consecutive tiles, chunks dealt round robin (monotonic in i)
  const uint64_t gmask = (1ull << glog) - 1ull;
  auto tile_of = [&](uint64_t i) { return (((uint64_t)blockIdx.x + (i >> glog) * gridDim.x) << glog) + (i & gmask); };
  uint64_t gfirst;
  unsigned int sfirst;
  thread_off(0, gfirst, sfirst);                    // the load layout does not depend on the tile
  uint64_t i = 0;
  for (uint64_t t = tile_of(0); t < ntiles; t = tile_of(++i)) {
    // physical index of the group's tile (all tile-index bits 0): the same for every thread of the group
    const uint64_t base = expand((t * NG + grp) << F, hs);
    double2 v[D];
    uint64_t gthread = base | gfirst;               // physical index of this thread's amplitudes with the register bits 0
    unsigned int sthread = sfirst;                  // its slot in the group's shared tile, register bits 0
    {
      // byte pointer + tabulated byte offsets, the pointer hidden from the optimiser: otherwise every address is
      // re-derived as psi + 16 (gthread + offset), four 32-bit instructions instead of two
      const char* src = reinterpret_cast<const char*>(psi + gthread);
      asm volatile("" : "+l"(src));
#pragma unroll
      for (int b = 0; b < D; ++b) {
        double2 a;     // explicit ld.global (the hidden pointer would be a generic LD otherwise); L1-allocating: 1 - 2 % faster than .cs
        asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(a.x), "=d"(a.y) : "l"(src + bt.gload[b]));
        v[b] = make_double2(g * a.x, g * a.y);      // the pending normalisation, once per pass
      }
    }
    // stage operands are fetched one stage ahead (s counts from 0 through the phases: uniform, so are the loads)
    int s = 0;
    unsigned int code;
    double2 a;
    unsigned long long qtab, mq;
    fetch(0, code, a, qtab, mq);
    for (int p = 0; p < nlayouts; ++p) {
      const int s_end = bt.s_end[p];
      for (; s < s_end; ++s) {
        unsigned int code_n;
        double2 a_n;
        unsigned long long qtab_n, mq_n;
        fetch(s + 1, code_n, a_n, qtab_n, mq_n);
        const Stage& st = bt.st[s];
        const unsigned long long tq = qtab ^ ((__popcll(gthread & mq) & 1) ? 0x8080808080808080ull : 0ull);
        double n0 = 0.0, n1 = 0.0;
        const unsigned int la = code >> 5, variant = code & 7u;
        if (la == 0) stage_by_variant<R, 0>(v, variant, st, a, tq, gthread, n0, n1);
        else if (la == 1) stage_by_variant<R, 1>(v, variant, st, a, tq, gthread, n0, n1);
        else if (R == 3 || la == 2) stage_by_variant<R, 2>(v, variant, st, a, tq, gthread, n0, n1);
        else stage_by_variant<R, R - 1>(v, variant, st, a, tq, gthread, n0, n1);
        if (code & 16u) sacc[s][tid] += n0 + n1;
        code = code_n;
        a = a_n;
        qtab = qtab_n;
        mq = mq_n;
      }
      if (p + 1 < nlayouts) {
        // transpose through the group's tile to the next layout
        const unsigned int s_old = sthread;
        uint64_t gn;
        thread_off(p + 1, gn, sthread);
        gthread = base | gn;
        const unsigned int s_new = sthread;
        group_sync();                             // every thread has read its slots of the previous transpose
#pragma unroll
        for (int b = 0; b < D; ++b) *reinterpret_cast<double2*>(tile + (s_old ^ bt.sbits[p][b])) = v[b];
        group_sync();
#pragma unroll
        for (int b = 0; b < D; ++b) v[b] = *reinterpret_cast<const double2*>(tile + (s_new ^ bt.sbits[p + 1][b]));
      } else {
        char* dst = reinterpret_cast<char*>(psi + gthread);
        asm volatile("" : "+l"(dst));
#pragma unroll
        for (int b = 0; b < D; ++b) st_amp(reinterpret_cast<double2*>(dst + bt.gstore[b]), v[b]);
      }
    }
  }
  double acc[MAX_STAGES];
#pragma unroll
  for (int k = 0; k < MAX_STAGES; ++k) acc[k] = sacc[k][tid];
  grid_finish<MAX_STAGES>(acc, partials, ctrl, res, seq, fin_mode, (double)bt.nstages, g, bt.nscale);
}

// ---------------------------------------------------------------------------
// The same batched pass with the CTA tile staged through shared memory by the bulk-copy engine (TMA, cp.async.bulk):
// the bytes in flight live in shared memory instead of registers, loads of tile i+1 and stores of tile i-1 overlap the
// FP64 work on tile i, and no thread computes a global address.
//   tile   = 2^H runs of 2^L physically contiguous amplitudes (L + H = 11): the low L physical bits entirely (batch axes
//            below bit L are simply part of the run) x the 2^H combinations of the batch axes at / above bit L.  One bulk
//            copy per run (>= 512 B), 2^H <= 64 copies per tile, issued by the lanes of warp 0 and tracked by an mbarrier
//            (loads) / a bulk group (stores).  Three tile buffers: tile i+1 loads while tile i is computed and tile
//            i-1 drains.
//   phases = as in k_fused_tile, but IN PLACE in the tile buffer: a thread reads the 2^3 amplitudes spanned by the three
//            register bits of the phase (any of the 11 tile-index bits), applies the stages of the phase, writes them back
//            to the same slots; one __syncthreads between phases.
// 96 KB of buffers + 16 KB of stage accumulators per CTA, 2 CTAs per SM.  Used when no dead bit lies below bit L.
// ---------------------------------------------------------------------------
struct TmaBatch {
  int nstages, nphases, L, H;
  unsigned char hbit[8];           // physical bit of tile-index bit L + i
  unsigned char lax[MAX_STAGES];   // register slot (0..2) of the axis of every stage inside its phase
  unsigned char sb[MAX_PHASES][4]; // register bits of every phase as tile-index bits, ascending
  unsigned char s_begin[MAX_PHASES], s_end[MAX_PHASES];
  double nscale[MAX_STAGES];
  Stage st[MAX_STAGES];
};
constexpr int TMA_BUFS = 3;

__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(TPB, 2) k_fused_tma(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                                     const __grid_constant__ TmaBatch bt, const uint64_t ntiles,
                                                     double* __restrict__ partials, Ctrl* ctrl, Res* res,
                                                     const unsigned long long seq, const int fin_mode) {
  extern __shared__ __align__(128) double2 tbuf[];    // TMA_BUFS x 2^TILE_LOG2 amplitudes
  __shared__ double sacc[MAX_STAGES][TPB];
  __shared__ __align__(8) unsigned long long full[TMA_BUFS];
  constexpr uint32_t TILE = 1u << TILE_LOG2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = bt.L, H = bt.H;
  const uint32_t run_bytes = 16u << L;
  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < TMA_BUFS; ++b) mbar_init(&full[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
#pragma unroll
  for (int k = 0; k < MAX_STAGES; ++k) sacc[k][tid] = 0.0;
  __syncthreads();
  const double g = ctrl->g;
  // physical offset of run r of a tile (lanes of warp 0 own runs lane, lane + 32)
  auto run_off = [&](int r) {
    uint64_t o = 0;
    for (int i = 0; i < H; ++i)
      if ((r >> i) & 1) o |= 1ull << bt.hbit[i];
    return o;
  };
  auto issue_load = [&](uint64_t t, int b) {           // whole warp 0
    const uint64_t base = expand(t << L, hs);
    if (lane == 0) mbar_arrive_expect_tx(&full[b], TILE * 16u);
    __syncwarp();
    for (int r = lane; r < (1 << H); r += 32)
      bulk_load(tbuf + (size_t)b * TILE + ((size_t)r << L), psi + base + run_off(r), run_bytes, &full[b]);
  };
  if (warp == 0 && blockIdx.x < ntiles) issue_load(blockIdx.x, 0);
  uint32_t it = 0;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int b = it % TMA_BUFS;
    const uint32_t parity = (it / TMA_BUFS) & 1u;
    const uint64_t tn = t + gridDim.x;
    if (warp == 0 && tn < ntiles) {
      bulk_wait_read<1>();                  // the stores that last read the next buffer (two tiles ago) have drained
      issue_load(tn, (it + 1) % TMA_BUFS);
    }
    double2* tb = tbuf + (size_t)b * TILE;
    const uint64_t base = expand(t << L, hs);
    mbar_wait(&full[b], parity);
    for (int p = 0; p < bt.nphases; ++p) {
      // this thread's tile index with the three register bits (ascending) cleared
      unsigned int sidx = tid;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const unsigned int lo = sidx & ((1u << bt.sb[p][k]) - 1u);
        sidx = ((sidx - lo) << 1) | lo;
      }
      uint64_t gthread = base | (uint64_t)(sidx & ((1u << L) - 1u));
      for (int i = 0; i < H; ++i)
        if ((sidx >> (L + i)) & 1u) gthread |= 1ull << bt.hbit[i];
      const unsigned int s0 = 1u << bt.sb[p][0], s1 = 1u << bt.sb[p][1], s2 = 1u << bt.sb[p][2];
      auto soff = [&](int bb) { return sidx | ((bb & 1) ? s0 : 0u) | ((bb & 2) ? s1 : 0u) | ((bb & 4) ? s2 : 0u); };
      double2 v[8];
#pragma unroll
      for (int bb = 0; bb < 8; ++bb) v[bb] = tb[soff(bb)];
      if (p == 0) {
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) { v[bb].x *= g; v[bb].y *= g; }   // the pending normalisation, once per pass
      }
      for (int s = bt.s_begin[p]; s < bt.s_end[p]; ++s) {
        const Stage& st = bt.st[s];
        double n0 = 0.0, n1 = 0.0;
        const int la = bt.lax[s];
        if (la == 0) stage_on_axis<3, 0>(v, st, gthread, n0, n1);
        else if (la == 1) stage_on_axis<3, 1>(v, st, gthread, n0, n1);
        else stage_on_axis<3, 2>(v, st, gthread, n0, n1);
        if (st.acc) sacc[s][tid] += n0 + n1;
      }
#pragma unroll
      for (int bb = 0; bb < 8; ++bb) tb[soff(bb)] = v[bb];
      if (p + 1 < bt.nphases) __syncthreads();
    }
    fence_proxy_async();                    // generic-proxy writes of the tile -> visible to the bulk-copy engine
    __syncthreads();
    if (warp == 0) {
      for (int r = lane; r < (1 << H); r += 32)
        bulk_store(psi + base + run_off(r), tb + ((size_t)r << L), run_bytes);
      bulk_commit();
    }
  }
  if (warp == 0) bulk_wait_all<0>();
  double acc[MAX_STAGES];
#pragma unroll
  for (int k = 0; k < MAX_STAGES; ++k) acc[k] = sacc[k][tid];
  grid_finish<MAX_STAGES>(acc, partials, ctrl, res, seq, fin_mode, (double)bt.nstages, g, bt.nscale);
}

// ---------------------------------------------------------------------------
// X / Z / C — 2x2 on axis p, in place.  Reads S, writes S.
// ---------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(TPB) k_apply1q(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                                 const int pbit, const double2 m00, const double2 m01,
                                                 const double2 m10, const double2 m11, const uint64_t npair) {
  const uint64_t chunk = (uint64_t)TPB * U;
  const uint64_t poff = 1ull << pbit;
  for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < npair; base += (uint64_t)gridDim.x * chunk) {
    double2 a[U], b[U];
    uint64_t s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      s[u] = expand(j, hs);
      if (j < npair) {
        a[u] = ld_amp(psi + s[u]);
        b[u] = ld_amp(psi + s[u] + poff);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      if (j < npair) {
        st_amp(psi + s[u], cmad2(m00, a[u], m01, b[u]));
        st_amp(psi + s[u] + poff, cmad2(m10, a[u], m11, b[u]));
      }
    }
  }
}

// Pending X byproducts made real: psi[i] <-> psi[i ^ fx] for every i whose bit b0 (the lowest bit of fx) is 0.
// One pass (reads S, writes S) for any number of flipped qubits.  `hs` = holes + {b0}.
template <int U>
__global__ void __launch_bounds__(TPB) k_xflip(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                               const uint64_t fx, const uint64_t npair) {
  const uint64_t chunk = (uint64_t)TPB * U;
  for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < npair; base += (uint64_t)gridDim.x * chunk) {
    double2 a[U], b[U];
    uint64_t s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      s[u] = expand(j, hs);
      if (j < npair) {
        a[u] = ld_amp(psi + s[u]);
        b[u] = ld_amp(psi + (s[u] ^ fx));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      if (j < npair) {
        st_amp(psi + s[u], b[u]);
        st_amp(psi + (s[u] ^ fx), a[u]);
      }
    }
  }
}

// diagonal 2x2 with m00 folded into the host scalar: psi[i|p] *= f.  Reads S/2, writes S/2.
template <int U>
__global__ void __launch_bounds__(TPB) k_scale_half(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                                    const int pbit, const double2 f, const uint64_t npair) {
  const uint64_t chunk = (uint64_t)TPB * U;
  const uint64_t poff = 1ull << pbit;
  for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < npair; base += (uint64_t)gridDim.x * chunk) {
    double2 b[U];
    uint64_t s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      s[u] = expand(j, hs) + poff;
      if (j < npair) b[u] = ld_amp(psi + s[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      if (j < npair) st_amp(psi + s[u], cmul(f, b[u]));
    }
  }
}

// scale every live amplitude (used when a pending complex scalar must be materialised)
template <int U>
__global__ void __launch_bounds__(TPB) k_scale_all(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                                   const double2 f, const uint64_t nlive) {
  const uint64_t chunk = (uint64_t)TPB * U;
  for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < nlive; base += (uint64_t)gridDim.x * chunk) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      if (j < nlive) {
        const uint64_t s = expand(j, hs);
        st_amp(psi + s, cmul(f, ld_amp(psi + s)));
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Reductions over pairs of axis p (read S, write nothing)
//   WHAT 0: <psi| op |psi>            -> (re, im)          expectation_single
//   WHAT 1: |row0 . pair|^2, |row1 . pair|^2 summed        half-norms of op|psi>
//   WHAT 2: lazy-mode probability of a rank-1 projector <c| on a qubit that still has queued CZs: with
//           t0 = c0 a, t1 = (-1)^{parity(x & mq)} c1 b it returns sum |t0 + t1|^2 and sum |t0 - t1|^2; the host
//           weights the two by the parity distribution of the not-yet-materialised CZ partners.
// ---------------------------------------------------------------------------
template <int U, int WHAT>
__global__ void __launch_bounds__(TPB) k_pair_reduce(const double2* __restrict__ psi,
                                                     const __grid_constant__ Holes hs, const int pbit,
                                                     const double2 m00, const double2 m01, const double2 m10,
                                                     const double2 m11, const uint64_t npair,
                                                     double* __restrict__ partials, Ctrl* ctrl, Res* res,
                                                     const unsigned long long seq, const double post,
                                                     const uint64_t mq) {
  const double g = ctrl->g;
  const uint64_t chunk = (uint64_t)TPB * U;
  const uint64_t poff = 1ull << pbit;
  double acc[2] = {0.0, 0.0};
  for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < npair; base += (uint64_t)gridDim.x * chunk) {
    double2 a[U], b[U];
    bool neg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      const uint64_t s = expand(j, hs);
      neg[u] = (WHAT == 2) && (__popcll(s & mq) & 1);
      if (j < npair) {
        a[u] = ld_amp(psi + s);
        b[u] = ld_amp(psi + s + poff);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      if (j < npair) {
        if (WHAT == 2) {
          const double2 t0 = cmul(m00, a[u]);
          double2 t1 = cmul(m01, b[u]);
          if (neg[u]) { t1.x = -t1.x; t1.y = -t1.y; }
          acc[0] += cnorm2(make_double2(t0.x + t1.x, t0.y + t1.y));
          acc[1] += cnorm2(make_double2(t0.x - t1.x, t0.y - t1.y));
          continue;
        }
        const double2 r0 = cmad2(m00, a[u], m01, b[u]);
        const double2 r1 = cmad2(m10, a[u], m11, b[u]);
        if (WHAT == 0) {
          // conj(a)*r0 + conj(b)*r1
          acc[0] += a[u].x * r0.x + a[u].y * r0.y + b[u].x * r1.x + b[u].y * r1.y;
          acc[1] += a[u].x * r0.y - a[u].y * r0.x + b[u].x * r1.y - b[u].y * r1.x;
        } else {
          acc[0] += cnorm2(r0);
          acc[1] += cnorm2(r1);
        }
      }
    }
  }
  grid_finish<2>(acc, partials, ctrl, res, seq, FIN_SCALED, post, g);
}

// sum |psi|^2 over the live sub-cube -> v[0]; (v[1] unused)
template <int U>
__global__ void __launch_bounds__(TPB) k_norm2(const double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                               const uint64_t nlive, double* __restrict__ partials, Ctrl* ctrl,
                                               Res* res, const unsigned long long seq, const double post) {
  const double g = ctrl->g;
  const uint64_t chunk = (uint64_t)TPB * U;
  double acc[2] = {0.0, 0.0};
  for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < nlive; base += (uint64_t)gridDim.x * chunk) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      if (j < nlive) acc[0] += cnorm2(ld_amp(psi + expand(j, hs)));
    }
  }
  grid_finish<2>(acc, partials, ctrl, res, seq, FIN_SCALED, post, g);
}

// <a|b> of two canonical (dense) vectors -> (re, im); no scaling
template <int U>
__global__ void __launch_bounds__(TPB) k_inner(const double2* __restrict__ x, const double2* __restrict__ y,
                                               const uint64_t n, double* __restrict__ partials, Ctrl* ctrl,
                                               Res* res, const unsigned long long seq) {
  const uint64_t chunk = (uint64_t)TPB * U;
  double acc[2] = {0.0, 0.0};
  for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < n; base += (uint64_t)gridDim.x * chunk) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t j = base + (uint64_t)u * TPB + threadIdx.x;
      if (j < n) {
        const double2 a = ld_amp(x + j), b = ld_amp(y + j);
        acc[0] += a.x * b.x + a.y * b.y;
        acc[1] += a.x * b.y - a.y * b.x;
      }
    }
  }
  grid_finish<2>(acc, partials, ctrl, res, seq, FIN_SCALED, 1.0, 1.0);
}

// ---------------------------------------------------------------------------
// finalize / flatten: out[i] = G * psi[perm(i)], where the bits of the
// canonical index i (qubit 0 = MSB) are scattered to their physical bits via
// byte-indexed lookup tables (5 tables x 256 entries cover 40 bits).
// ---------------------------------------------------------------------------
struct GatherTables {
  uint64_t t[5][256];
};
__global__ void __launch_bounds__(TPB) k_gather(const double2* __restrict__ psi, double2* __restrict__ out,
                                                const GatherTables* __restrict__ tabs, const uint64_t first,
                                                const uint64_t count, const double2 h, const Ctrl* ctrl,
                                                const uint64_t fx, const uint64_t fz) {
  // fx / fz: pending Pauli frame (physical-bit masks): true[i] = (-1)^{parity(p(i) & fz)} stored[p(i) ^ fx]
  __shared__ uint64_t st[5][256];
  for (int i = threadIdx.x; i < 5 * 256; i += TPB) st[i >> 8][i & 255] = tabs->t[i >> 8][i & 255];
  __syncthreads();
  const double g = ctrl->g;
  const double2 f = make_double2(h.x * g, h.y * g);
  for (uint64_t k = (uint64_t)blockIdx.x * TPB + threadIdx.x; k < count; k += (uint64_t)gridDim.x * TPB) {
    const uint64_t i = first + k;
    const uint64_t p = st[0][i & 255] | st[1][(i >> 8) & 255] | st[2][(i >> 16) & 255] | st[3][(i >> 24) & 255] |
                       st[4][(i >> 32) & 255];
    double2 a = cmul(f, __ldg(psi + (p ^ fx)));
    if (__popcll(p & fz) & 1) { a.x = -a.x; a.y = -a.y; }
    out[k] = a;
  }
}

// <a|b> of two registers in their own physical layouts (no canonical copies): canonical index i is scattered to each
// layout through that register's byte tables.  -> (re, im), unscaled.
__global__ void __launch_bounds__(TPB) k_inner_gather(const double2* __restrict__ xa, const GatherTables* __restrict__ ta,
                                                      const double2* __restrict__ xb, const GatherTables* __restrict__ tb,
                                                      const uint64_t n, double* __restrict__ partials, Ctrl* ctrl,
                                                      Res* res, const unsigned long long seq, const uint64_t fxa,
                                                      const uint64_t fza, const uint64_t fxb, const uint64_t fzb) {
  __shared__ uint64_t sa[5][256];
  __shared__ uint64_t sb[5][256];
  for (int i = threadIdx.x; i < 5 * 256; i += TPB) {
    sa[i >> 8][i & 255] = ta->t[i >> 8][i & 255];
    sb[i >> 8][i & 255] = tb->t[i >> 8][i & 255];
  }
  __syncthreads();
  double acc[2] = {0.0, 0.0};
  for (uint64_t i = (uint64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (uint64_t)gridDim.x * TPB) {
    const uint64_t pa = sa[0][i & 255] | sa[1][(i >> 8) & 255] | sa[2][(i >> 16) & 255] | sa[3][(i >> 24) & 255] |
                        sa[4][(i >> 32) & 255];
    const uint64_t pb = sb[0][i & 255] | sb[1][(i >> 8) & 255] | sb[2][(i >> 16) & 255] | sb[3][(i >> 24) & 255] |
                        sb[4][(i >> 32) & 255];
    const double2 a = __ldg(xa + (pa ^ fxa));
    double2 b = __ldg(xb + (pb ^ fxb));
    if ((__popcll(pa & fza) ^ __popcll(pb & fzb)) & 1) { b.x = -b.x; b.y = -b.y; }   // pending Pauli frames of both sides
    acc[0] += a.x * b.x + a.y * b.y;
    acc[1] += a.x * b.y - a.y * b.x;
  }
  grid_finish<2>(acc, partials, ctrl, res, seq, FIN_SCALED, 1.0, 1.0);
}

// Sparse read-out (Statevector.to_dict / to_prob_dict, statevec.py:591-678): canonical indices and amplitudes of the
// entries with |amp| > atol are appended to (idx, amp) through one atomic counter; at most `cap` are stored, the
// counter keeps counting so that the host can tell an overflow.
__global__ void __launch_bounds__(TPB) k_nonzero(const double2* __restrict__ psi, const GatherTables* __restrict__ tabs,
                                                 const uint64_t count, const double2 h, const Ctrl* ctrl,
                                                 const double atol, unsigned long long* __restrict__ counter,
                                                 unsigned long long* __restrict__ idx_out, double2* __restrict__ amp_out,
                                                 const unsigned long long cap, const uint64_t fx, const uint64_t fz) {
  __shared__ uint64_t st[5][256];
  for (int i = threadIdx.x; i < 5 * 256; i += TPB) st[i >> 8][i & 255] = tabs->t[i >> 8][i & 255];
  __syncthreads();
  const double g = ctrl->g;
  const double2 f = make_double2(h.x * g, h.y * g);
  for (uint64_t i = (uint64_t)blockIdx.x * TPB + threadIdx.x; i < count; i += (uint64_t)gridDim.x * TPB) {
    const uint64_t p = st[0][i & 255] | st[1][(i >> 8) & 255] | st[2][(i >> 16) & 255] | st[3][(i >> 24) & 255] |
                       st[4][(i >> 32) & 255];
    double2 a = cmul(f, __ldg(psi + (p ^ fx)));
    if (__popcll(p & fz) & 1) { a.x = -a.x; a.y = -a.y; }
    if (sqrt(cnorm2(a)) > atol) {
      const unsigned long long slot = atomicAdd(counter, 1ull);
      if (slot < cap) {
        idx_out[slot] = i;
        amp_out[slot] = a;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// k-qubit dense operator (k = 2, 3, 4) on arbitrary axes, in place: each thread owns one group of
// 2^K amplitudes.  Used by Statevector.evolve (statevec.py:361-395) and by the density-matrix path,
// where a single-qubit channel is a 2-axis (row bit, column bit) superoperator and a two-qubit
// channel a 4-axis one.  `offs[b]` = physical offset contributed by operator-index b (precomputed on
// the host: bit (K-1-i) of b <-> target qubit i).  The matrix sits in shared memory (row-major).
// Reads S, writes S; 2^K complex FMAs per amplitude (K = 4 is FP64-bound, not HBM-bound).
// ---------------------------------------------------------------------------
struct GroupOffsets {
  unsigned long long off[16];
};
template <int K>
__global__ void __launch_bounds__(TPB) k_evolve(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                                const __grid_constant__ GroupOffsets go,
                                                const double2* __restrict__ op, const uint64_t ngroups) {
  constexpr int D = 1 << K;
  __shared__ double2 m[D * D];
  for (int i = threadIdx.x; i < D * D; i += TPB) m[i] = op[i];
  __syncthreads();
  for (uint64_t j = (uint64_t)blockIdx.x * TPB + threadIdx.x; j < ngroups; j += (uint64_t)gridDim.x * TPB) {
    const uint64_t base = expand(j, hs);
    double2 v[D];
#pragma unroll
    for (int b = 0; b < D; ++b) v[b] = ld_amp(psi + base + go.off[b]);
#pragma unroll
    for (int a = 0; a < D; ++a) {
      double re = 0.0, im = 0.0;
#pragma unroll
      for (int b = 0; b < D; ++b) {
        const double2 c = m[a * D + b];
        re = fma(c.x, v[b].x, re);
        re = fma(-c.y, v[b].y, re);
        im = fma(c.x, v[b].y, im);
        im = fma(c.y, v[b].x, im);
      }
      st_amp(psi + base + go.off[a], make_double2(re, im));
    }
  }
}

// ---------------------------------------------------------------------------
// Density matrices are stored vectorised (|rho>> over row and column qubits).  Traces only touch the
// 2^n "diagonal" amplitudes rho[x, x]: x is scattered to its row and column bits through byte tables.
//   acc[0..1] = sum_x sum_ab m[a][b] * rho[(x, t=b), (x, t=a)]   = Tr(op_t rho)      (skipped if !has_target)
//   acc[2..3] = sum_x sum_a rho[(x, a), (x, a)]                  = Tr(rho)
// density_matrix.py:263-290 (expectation_single), :354-365 (normalize).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_dm_diag(const double2* __restrict__ psi, const GatherTables* __restrict__ tabs,
                                                 const uint64_t nx, const int has_target, const uint64_t roff,
                                                 const uint64_t coff, const double2 m00, const double2 m01,
                                                 const double2 m10, const double2 m11, double* __restrict__ partials,
                                                 Ctrl* ctrl, Res* res, const unsigned long long seq) {
  __shared__ uint64_t st[5][256];
  for (int i = threadIdx.x; i < 5 * 256; i += TPB) st[i >> 8][i & 255] = tabs->t[i >> 8][i & 255];
  __syncthreads();
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (uint64_t x = (uint64_t)blockIdx.x * TPB + threadIdx.x; x < nx; x += (uint64_t)gridDim.x * TPB) {
    const uint64_t base = st[0][x & 255] | st[1][(x >> 8) & 255] | st[2][(x >> 16) & 255] | st[3][(x >> 24) & 255] |
                          st[4][(x >> 32) & 255];
    const double2 r00 = __ldg(psi + base);
    if (has_target) {
      const double2 r01 = __ldg(psi + base + coff);          // row t = 0, col t = 1
      const double2 r10 = __ldg(psi + base + roff);          // row t = 1, col t = 0
      const double2 r11 = __ldg(psi + base + roff + coff);
      // Tr(op rho) = sum_ab op[a][b] rho[b][a]
      const double2 t0 = cmad2(m00, r00, m01, r10);
      const double2 t1 = cmad2(m10, r01, m11, r11);
      acc[0] += t0.x + t1.x;
      acc[1] += t0.y + t1.y;
      acc[2] += r00.x + r11.x;
      acc[3] += r00.y + r11.y;
    } else {
      acc[2] += r00.x;
      acc[3] += r00.y;
    }
  }
  grid_finish<4>(acc, partials, ctrl, res, seq, FIN_SCALED, 1.0, 1.0);
}

// ---------------------------------------------------------------------------
// Sharded states: the half of the live sub-cube with bit p = bitval, in dense rank order, is
// packed into / unpacked from a contiguous staging buffer for the NVLink half-shard exchange
// that swaps a global (rank-index) qubit with the local qubit on bit p.  `hs` = holes + {p}.
// ---------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(TPB) k_pack_half(const double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                                   const uint64_t pbitval, const uint64_t first, const uint64_t count,
                                                   double2* __restrict__ out) {
  const uint64_t chunk = (uint64_t)TPB * U;
  for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < count; base += (uint64_t)gridDim.x * chunk) {
    double2 a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t k = base + (uint64_t)u * TPB + threadIdx.x;
      if (k < count) a[u] = ld_amp(psi + (expand(first + k, hs) | pbitval));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t k = base + (uint64_t)u * TPB + threadIdx.x;
      if (k < count) out[k] = a[u];
    }
  }
}
template <int U>
__global__ void __launch_bounds__(TPB) k_unpack_half(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                                     const uint64_t pbitval, const uint64_t first, const uint64_t count,
                                                     const double2* __restrict__ in, const double2 f) {
  const uint64_t chunk = (uint64_t)TPB * U;
  for (uint64_t base = (uint64_t)blockIdx.x * chunk; base < count; base += (uint64_t)gridDim.x * chunk) {
    double2 a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t k = base + (uint64_t)u * TPB + threadIdx.x;
      if (k < count) a[u] = in[k];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t k = base + (uint64_t)u * TPB + threadIdx.x;
      if (k < count) st_amp(psi + (expand(first + k, hs) | pbitval), cmul(f, a[u]));
    }
  }
}

// ---------------------------------------------------------------------------
// Sharded states, peer-memory exchange (one process per GPU, the partner's register mapped through CUDA IPC over
// NVLink / NVSwitch).  Swapping the rank-index qubit on rank bit k with the local qubit on physical bit p exchanges, between
// the two ranks that differ in bit k, the half with bit p = 1 - B of the rank whose bit k is B with the half with bit
// p = B of its partner, element by element:
//     mine[e(i) | my_or]  <->  peer[e(i) | peer_or]          (each scaled by the ratio of the ranks' bookkeeping scalars)
// Each rank runs this kernel over ITS half of the index range [first, first + count), so every pair is touched by exactly
// one thread of one GPU: in place on both sides, no staging buffers, no pack / unpack passes, and the NVLink traffic of
// a rank is count * 16 B of loads plus count * 16 B of stores in each direction.  The host brackets the launch with two
// stream-ordered cross-rank barriers (graphix_b200/sharded.py).
// ---------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(TPB) k_swap_peer(double2* __restrict__ mine, double2* __restrict__ peer,
                                                   const __grid_constant__ Holes hs, const uint64_t my_or,
                                                   const uint64_t peer_or, const uint64_t first, const uint64_t count,
                                                   const double2 f_mine, const double2 f_peer, const int scale,
                                                   const int mode, const uint64_t span) {
  // mode 0 = swap; 1 / 2 = diagnostics only (tools/microbench_swap.py): pull (mine <- peer) / push (peer <- mine).
  // span > 0: every CTA walks ONE contiguous range of `span` elements (it stays inside the same 2 MB pages of the local
  // and of the peer mapping for many iterations); span = 0: plain grid stride.
  const uint64_t chunk = (uint64_t)TPB * U;
  const uint64_t lo = span ? (uint64_t)blockIdx.x * span : (uint64_t)blockIdx.x * chunk;
  const uint64_t hi = span ? (lo + span < count ? lo + span : count) : count;
  const uint64_t step = span ? chunk : (uint64_t)gridDim.x * chunk;
  for (uint64_t base = lo; base < hi; base += step) {
    double2 a[U], b[U];
    uint64_t e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t k = base + (uint64_t)u * TPB + threadIdx.x;
      e[u] = expand(first + k, hs);
      if (k < hi) {
        if (mode != 2) b[u] = ld_amp(peer + (e[u] | peer_or));      // NVLink load
        if (mode != 1) a[u] = ld_amp(mine + (e[u] | my_or));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t k = base + (uint64_t)u * TPB + threadIdx.x;
      if (k < hi) {
        if (mode != 1) st_amp(peer + (e[u] | peer_or), scale ? cmul(f_peer, a[u]) : a[u]);   // NVLink store
        if (mode != 2) st_amp(mine + (e[u] | my_or), scale ? cmul(f_mine, b[u]) : b[u]);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// general add_nodes: psi[i | spread(j)] = v[j] * psi[i] for j = 1 .. 2^k-1
// (the j = 0 term is applied afterwards with k_scale_all).  `newbits` lists the
// physical bit of each new qubit, MSB (first new qubit) first.
// ---------------------------------------------------------------------------
struct NewBits {
  int k;
  unsigned char pos[32];
};
__global__ void __launch_bounds__(TPB) k_tensor(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                                const __grid_constant__ NewBits nb, const double2* __restrict__ vec,
                                                const uint64_t nlive, const int nlive_bits) {
  const uint64_t total = nlive << nb.k;
  for (uint64_t t = (uint64_t)blockIdx.x * TPB + threadIdx.x; t < total; t += (uint64_t)gridDim.x * TPB) {
    const uint64_t j = t >> nlive_bits;
    if (j == 0) continue;
    const uint64_t i = t & (nlive - 1);
    uint64_t d = expand(i, hs);
    const double2 a = psi[d];
    for (int b = 0; b < nb.k; ++b) d |= ((j >> (nb.k - 1 - b)) & 1ull) << nb.pos[b];
    psi[d] = cmul(vec[j], a);
  }
}

}  // namespace gxk
