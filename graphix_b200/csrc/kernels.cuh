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

struct Res {             // one slot of the host-mapped result ring
  double v[8];
  unsigned long long seq;
  unsigned long long pad[7];
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
__device__ __forceinline__ double2 cmad2(const double2 a, const double2 x, const double2 b, const double2 y) {
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
                                            Res* res, unsigned long long seq, int mode, double post, double g) {
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
      ctrl->g = 1.0 / sqrt(tot[0]);
      res->v[0] = tot[0];
    } else if (mode == FIN_CONTRACT_DIST) {
      res->v[0] = tot[0];
    } else if (mode == FIN_BATCH) {
#pragma unroll
      for (int k = 0; k < NV; ++k) res->v[k] = tot[k];
      ctrl->g = 1.0 / sqrt(tot[(int)post - 1]);
    } else if (mode == FIN_ONESHOT) {
      // sharded, non-blocking: publish the local norms; the pending factor g has been consumed by this pass and
      // stays 1 until the host sets it again from the all-reduced norm (gxsv_set_norm)
#pragma unroll
      for (int k = 0; k < NV; ++k) res->v[k] = tot[k];
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
                                               const unsigned long long seq, const int fin_mode) {
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
  grid_finish<1>(acc, partials, ctrl, res, seq, fin_mode, 1.0, g);
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
constexpr int MAX_STAGES = 8;
struct Stage {
  double2 m00, m01, m10, m11;
  unsigned long long mq, mv;
  int axis;
  int pad;
};
struct Batch {
  int nstages;
  int naxes;
  unsigned char axbit[8];
  Stage st[MAX_STAGES];
};
template <int A>
__global__ void __launch_bounds__(TPB, (A <= 3) ? 4 : 2) k_fused_batch(double2* __restrict__ psi, const __grid_constant__ Holes hs,
                                                     const __grid_constant__ Batch bt, const uint64_t ngroups,
                                                     double* __restrict__ partials, Ctrl* ctrl, Res* res,
                                                     const unsigned long long seq, const int fin_mode) {
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
  // per-thread stage accumulators live in shared memory (one column per thread, no conflicts): 16 registers less,
  // which is what gets the kernel to 4 CTAs per SM
  __shared__ double sacc[MAX_STAGES][TPB];
#pragma unroll
  for (int k = 0; k < MAX_STAGES; ++k) sacc[k][threadIdx.x] = 0.0;
  for (uint64_t j = (uint64_t)blockIdx.x * TPB + threadIdx.x; j < ngroups; j += (uint64_t)gridDim.x * TPB) {
    const uint64_t base = expand(j, hs);
    double2 v[D];
#pragma unroll
    for (int b = 0; b < D; ++b) {
      const double2 a = ld_amp(psi + base + off_of(b));
      v[b] = make_double2(g * a.x, g * a.y);       // the pending normalisation, once per pass
    }
#pragma unroll
    for (int s = 0; s < MAX_STAGES; ++s) {
      if (s < bt.nstages) {
        const Stage& st = bt.st[s];
        double nrm = 0.0;
#pragma unroll
        for (int a = 0; a < A; ++a) {
          if (st.axis == a) {
#pragma unroll
            for (int lo = 0; lo < D; ++lo) {
              if ((lo >> a) & 1) continue;
              const int hi = lo | (1 << a);
              const uint64_t idx = base + off_of(lo);
              // queued CZ signs: conditional negation, no multiplies
              const bool nq = __popcll(idx & st.mq) & 1;
              const bool nv = __popcll(idx & st.mv) & 1;
              const double2 t1 = nq ? make_double2(-v[hi].x, -v[hi].y) : v[hi];
              const double2 r0 = cmad2(st.m00, v[lo], st.m01, t1);
              double2 r1 = cmad2(st.m10, v[lo], st.m11, t1);
              if (nv) { r1.x = -r1.x; r1.y = -r1.y; }
              nrm += cnorm2(r0) + cnorm2(r1);
              v[lo] = r0;
              v[hi] = r1;
            }
          }
        }
        sacc[s][threadIdx.x] += nrm;
      }
    }
#pragma unroll
    for (int b = 0; b < D; ++b) st_amp(psi + base + off_of(b), v[b]);
  }
  double acc[MAX_STAGES];
#pragma unroll
  for (int k = 0; k < MAX_STAGES; ++k) acc[k] = sacc[k][threadIdx.x];
  grid_finish<MAX_STAGES>(acc, partials, ctrl, res, seq, fin_mode, (double)bt.nstages, g);
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
                                                const uint64_t count, const double2 h, const Ctrl* ctrl) {
  __shared__ uint64_t st[5][256];
  for (int i = threadIdx.x; i < 5 * 256; i += TPB) st[i >> 8][i & 255] = tabs->t[i >> 8][i & 255];
  __syncthreads();
  const double g = ctrl->g;
  const double2 f = make_double2(h.x * g, h.y * g);
  for (uint64_t k = (uint64_t)blockIdx.x * TPB + threadIdx.x; k < count; k += (uint64_t)gridDim.x * TPB) {
    const uint64_t i = first + k;
    const uint64_t p = st[0][i & 255] | st[1][(i >> 8) & 255] | st[2][(i >> 16) & 255] | st[3][(i >> 24) & 255] |
                       st[4][(i >> 32) & 255];
    out[k] = cmul(f, __ldg(psi + p));
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
