// gxsv.cu — host side of the gxsv engine and its C ABI (include/gxsv.h).
//
// Holds the private physical layout (which physical bit carries which logical
// qubit, which bits are dead), the pending real/complex normalisation, the
// queue of un-issued CZs, and launches the kernels of kernels.cuh.
//
// Reference semantics reproduced (TeamGraphix/graphix):
//   graphix/sim/statevec.py:52-750   class Statevector
//   graphix/sim/statevec.py:841-1198 numba kernels
#include "../../include/gxsv.h"
#include "kernels.cuh"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

using namespace gxk;
typedef std::complex<double> cplx;

namespace {

constexpr int HOLE_MIN = 8;     // dead bits are only ever created at physical bit >= HOLE_MIN
constexpr int TILE_BITS = 11;   // k_contract<4,true>: a CTA tile spans the lowest 11 live bits
constexpr int U_DEF = 4;
constexpr int RING = 4096;    // result slots (256 B each, pinned + mapped)
constexpr int GXSV_VERSION = 100;
constexpr int MAX_BATCH_AXES = 7;  // distinct physical bits one batched pass may act on (GXSV_BATCH_AXES = 1..7).
// Up to 3 axes run in k_fused_batch (2^3 amplitudes per thread in registers, 4 CTAs per SM); 4..6 axes run in
// k_fused_tile, which keeps three axes at a time in registers and transposes the CTA tile through shared memory
// between phases.  (Round 1 measured a 4-axis all-register pass: 128 registers, 2 CTAs per SM, 0.64 of the copy rate.)
constexpr int DEFAULT_CHUNK_LOG2 = 3;   // consecutive 256-group blocks / tiles a CTA takes at a time (GXSV_CHUNK_LOG2)

struct Qubit {
  int id;       // stable identity (queued CZs refer to ids: logical indices shift, tiles relocate bits)
  int phys;     // physical bit, -1 while virtual (lazy mode: prepared but not yet materialised)
  cplx s0, s1;  // single-qubit state while virtual
};

struct ProfRec {
  int kind;
  cudaEvent_t e0, e1;
};

// A projection whose norm has not been looked at yet (non-strict mode).  The reference keeps half 0 of the
// projected pair unless its norm is <= 1e-10 (statevec.py:1187-1190); when the larger row of the projector is row 1
// the global phase of the result depends on that choice.  It is guessed when the projection is issued
// (`assumed`) and corrected when the norm is finally known (`ratio` = |row0|^2 / |row1|^2, `phi` = the unit phase).
struct StageNote {
  double an = 0.0;       // > 0: the norm^2 after this stage is an * (norm^2 before it), known without looking (Stage::acc = 0)
  double ratio = 0.0;
  cplx phi = 1.0;
  bool big1 = false;
  bool assumed = false;
};
struct Deferred {
  unsigned long long first = 0;   // result slot
  int second = 0;                 // stages
  StageNote note[MAX_STAGES];
};

}  // namespace

struct gxsv {
  int device = 0;
  cudaStream_t stream = nullptr;
  double2* psi = nullptr;
  bool own_buf = true;
  int cap_bits = 0;
  int phys_bits = 0;       // all live physical bits are < phys_bits
  uint64_t live_mask = 0;  // set bits = physical bits carrying a materialised qubit
  std::vector<Qubit> q;    // logical order (reference order: index 0 = MSB)
  std::vector<std::pair<int, int>> pend;  // queued CZs as (qubit id, qubit id)
  int next_id = 0;
  cplx hscale = 1.0;       // host part of the pending scalar: true psi = g(device) * hscale * stored
  // device control + result ring
  Ctrl* ctrl = nullptr;
  double* partials = nullptr;
  Res* ring_host = nullptr;
  Res* ring_dev = nullptr;
  unsigned long long seq = 0;
  GatherTables* tabs_dev = nullptr;
  double2* op_dev = nullptr;   // 16x16 operator scratch for gxsv_evolve
  double2* stage[2] = {nullptr, nullptr};
  size_t stage_amps = 0;
  // options
  int fusion = 1;
  int strict = 1;
  int profile = 0;
  int batch_axes = MAX_BATCH_AXES;   // distinct physical bits per batched pass (GXSV_BATCH_AXES)
  int chunk_log2 = DEFAULT_CHUNK_LOG2;
  int max_phases = MAX_PHASES;   // register phases one batched pass may need (GXSV_MAX_PHASES = 1..4)
  int dist_queue = 0;      // sharded strict mode: the orchestrator accepts queued (analytic) projections, see GXSV_OPT_DIST_QUEUE
  int tile_regs = 4;       // register axes per phase of k_fused_tile (GXSV_TILE_REGS = 3 | 4; measured on config 2: 4 ->
                           // 0.378 ms per pass / 44.8 k commands/s, 3 -> 0.518 ms / 32.9 k)
  int analytic_norms = 1;  // skip the norm accumulation of stages whose outcome probability is known to be 1/2
  int use_tma = 0;         // 1: batched passes stage CTA tiles through shared memory with bulk copies (k_fused_tma; measured slower for many
                           // stages per pass, see kernels.cuh) when the layout allows
  int stage_prefetch = 1;         // k_fused_tile fetches the operands of stage s + 1 from shared memory while stage s runs (GXSV_STAGE_PREFETCH)
  int tile_warps = 0;             // warps sharing one tile of k_fused_tile: 0 = by the number of axes, else 1 / 4 / 8 (GXSV_TILE_WARPS)
  int tile_io = 1;                // k_fused_tile: extra load / store layouts over the lowest tile bits: 0 never, 1 when the three
                                  // lowest tile bits are all batch axes (default), 2 always (GXSV_TILE_IO)
  bool defer_partner_cz = true;   // batched stages leave the partner's CZs queued (GXSV_DEFER_VCZ=0: apply them in the stage)
  int batch_max = 0;       // > 0: queue up to this many fused projections and run them in one pass (non-strict only)
  Batch batch;             // the open batch (nstages == 0: none)
  std::vector<Deferred> deferred;   // projections awaiting their norm check (non-strict mode)
  StageNote open_notes[MAX_STAGES]; // notes of the stages in the open batch
  int dist = 0;            // sharded: projections publish local norms, the host sets g after the all-reduce
  int sm_count = 148;
  gxsv_stats_t stats;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  std::string err;
  // peer registers of the ranks this shard exchanges with (CUDA IPC mappings, one slot per rank-index bit)
  // Pauli frame of pending X / Z byproducts on materialised qubits (physical-bit masks, lazy mode only):
  // true state = prod_q Z_q^{fz_q} X_q^{fx_q} applied to what the register (plus its queued work) holds.  X and Z at the
  // end of a pattern then cost no pass: read-outs apply the frame on the fly, anything else makes it real first.
  uint64_t fx = 0, fz = 0;
  bool lazy_frame = true;
  bool attr_tile = false, attr_tma = false;   // dynamic shared-memory opt-in done for this handle's device
  double2* peer_psi[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double2* exported_psi = nullptr;
};

static thread_local std::string g_err;

namespace {

int fail(gxsv_t* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  g_err = buf;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      cudaGetLastError(); /* clear the (non-sticky) error so that later launch checks start clean */ \
      return fail(h, e_ == cudaErrorMemoryAllocation ? GXSV_ENOMEM : GXSV_ECUDA, "%s: %s", #call, \
                  cudaGetErrorString(e_));                                                         \
    }                                                                                              \
  } while (0)

inline double2 d2(cplx z) { return make_double2(z.real(), z.imag()); }

// sorted list of dead bits below phys_bits, plus up to two extra (axis) bits
Holes make_holes(const gxsv_t* h, int extra0 = -1, int extra1 = -1) {
  Holes r;
  r.n = 0;
  for (int b = 0; b < h->phys_bits; ++b) {
    const bool dead = !((h->live_mask >> b) & 1ull);
    if (dead || b == extra0 || b == extra1) r.pos[r.n++] = (unsigned char)b;
  }
  // an extra bit at/above phys_bits needs no insertion (rank indices have no bits there)
  return r;
}

Holes holes_of_mask(uint64_t live, int extent, int extra0 = -1) {
  Holes r;
  r.n = 0;
  for (int b = 0; b < extent; ++b) {
    const bool dead = !((live >> b) & 1ull);
    if (dead || b == extra0) r.pos[r.n++] = (unsigned char)b;
  }
  return r;
}

int nlive_bits(const gxsv_t* h) { return __builtin_popcountll(h->live_mask); }

int grid_for(const gxsv_t* h, uint64_t work, int per_thread) {
  const uint64_t chunk = (uint64_t)TPB * per_thread;
  uint64_t blocks = (work + chunk - 1) / chunk;
  const uint64_t cap = (uint64_t)h->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// --- profiling brackets ---------------------------------------------------
cudaEvent_t get_event(gxsv_t* h) {
  if (!h->ev_pool.empty()) {
    cudaEvent_t e = h->ev_pool.back();
    h->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

int flush_batch(gxsv_t* h);
int settle_frame(gxsv_t* h);
thread_local TmaBatch g_tma;        // scratch of flush_batch (kernel parameters are copied at launch)
thread_local Holes g_tma_holes;

struct Launch {
  gxsv_t* h;
  int kind;
  cudaEvent_t e0 = nullptr;
  Launch(gxsv_t* h_, int kind_, double bytes, bool frame_aware = false) : h(h_), kind(kind_) {
    if (h->batch.nstages) flush_batch(h);   // queued projections go first: everything else sees their result
    if (!frame_aware && (h->fx | h->fz)) settle_frame(h);   // then the pending X / Z byproducts, made real
    h->stats.launches[kind] += 1;
    h->stats.algo_bytes[kind] += bytes;
    if (h->profile) {
      e0 = get_event(h);
      cudaEventRecord(e0, h->stream);
    }
  }
  ~Launch() {
    if (h->profile) {
      cudaEvent_t e1 = get_event(h);
      cudaEventRecord(e1, h->stream);
      h->prof.push_back({kind, e0, e1});
    }
  }
};

void prof_collect(gxsv_t* h) {
  if (h->prof.empty()) return;
  cudaStreamSynchronize(h->stream);
  for (auto& r : h->prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) h->stats.device_ms[r.kind] += ms;
    h->ev_pool.push_back(r.e0);
    h->ev_pool.push_back(r.e1);
  }
  h->prof.clear();
}

// --- result ring ------------------------------------------------------------
Res* next_slot(gxsv_t* h, unsigned long long* seq_out) {
  h->seq += 1;
  *seq_out = h->seq;
  return h->ring_dev + (h->seq % RING);
}

int wait_slot(gxsv_t* h, unsigned long long seq, double out[RES_VALUES]) {
  volatile Res* r = h->ring_host + (seq % RING);
  unsigned long spins = 0;
  while (r->seq != seq) {
    if ((++spins & 0x3fff) == 0) {
      cudaError_t e = cudaStreamQuery(h->stream);
      if (e == cudaSuccess) {
        if (r->seq == seq) break;
        // kernel finished without publishing: should not happen
        if (spins > (1ul << 26)) return fail(h, GXSV_ECUDA, "result slot %llu never published", seq);
      } else if (e != cudaErrorNotReady) {
        return fail(h, GXSV_ECUDA, "device error while waiting: %s", cudaGetErrorString(e));
      }
    }
  }
  __sync_synchronize();
  for (int i = 0; i < RES_VALUES; ++i) out[i] = r->v[i];
  return GXSV_OK;
}

int check_launch(gxsv_t* h, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, GXSV_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return GXSV_OK;
}

// Sign table of a stage acting on register axis `la` of `nb` register axes (physical bits regbit[]): byte k = 0x80 iff
// the k-th amplitude pair (slot lo = k with a zero bit inserted at position la) has odd parity of
// (physical bits of the register axes set in lo) & mask.
unsigned long long parity_table(uint64_t mask, const unsigned char* regbit, int nb, int la) {
  unsigned long long t = 0;
  for (int k = 0; k < (1 << (nb - 1)); ++k) {
    const int lo = ((k >> la) << (la + 1)) | (k & ((1 << la) - 1));
    int par = 0;
    for (int i = 0; i < nb; ++i)
      if (((lo >> i) & 1) && ((mask >> regbit[i]) & 1ull)) par ^= 1;
    if (par) t |= 0x80ull << (8 * k);
  }
  return t;
}

// Greedy phase plan of a batch on more than three axes: every phase is the longest run of consecutive stages that
// acts on at most R distinct batch axes (those are held in registers, k_fused_tile).  Returns the number of phases
// (out may be NULL: the queueing code only asks whether one more stage still fits MAX_PHASES).
// With `out` (then live_mask / tile_log2 / io_mode are needed) the LAYOUTS of the pass are tabulated (kernels.cuh,
// k_fused_tile): the tile-index bits are the A axis bits and the F = tile_log2 - A lowest live non-axis bits; a layout
// keeps R of them in registers and gives each of the others a bit of the thread's index in its group.
//   * a phase keeps its (padded) axes in registers; its thread bits are ordered so that the three lowest lane bits land
//     in different bank groups of the (XOR-folded) shared tile whenever possible;
//   * the first / last layout is what the tile is loaded / stored in.  io_mode 1: when the three lowest tile bits are
//     all batch axes, an extra stage-less layout is put in front / at the end whose thread bits are the LOWEST tile-index
//     bits in physical order and whose register bits are the R highest — unless the first / last phase happens to
//     have exactly those register bits; io_mode 0: never, 2: always.
int plan_phases(const Batch& bt, TileBatch* out, int R, uint64_t live_mask = 0, int tile_log2 = 0, int io_mode = 1) {
  struct Phase { int T[4]; int s_begin, s_end; };
  Phase ph[MAX_STAGES];
  int nph = 0;
  int i = 0;
  const int A = bt.naxes;
  while (i < bt.nstages) {
    Phase& P = ph[nph];
    int nt = 0, j = i;
    for (; j < bt.nstages; ++j) {
      const int a = bt.st[j].axis;
      bool in = false;
      for (int k = 0; k < nt; ++k) in = in || (P.T[k] == a);
      if (in) continue;
      if (nt == R) break;
      P.T[nt++] = a;
    }
    // pad the register set to R axes with batch axes that the phase does not use, highest physical bits first (the
    // best chance to coincide with the load / store layout)
    while (nt < R) {
      int best = -1;
      for (int a = 0; a < A; ++a) {
        bool in = false;
        for (int k = 0; k < nt; ++k) in = in || (P.T[k] == a);
        if (!in && (best < 0 || bt.axbit[a] > bt.axbit[best])) best = a;
      }
      if (best < 0) break;      // fewer axes than register bits cannot happen (A >= R on this path)
      P.T[nt++] = best;
    }
    std::sort(P.T, P.T + nt);
    P.s_begin = i;
    P.s_end = j;
    nph += 1;
    i = j;
  }
  if (!out || nph > MAX_PHASES) return nph;

  const int WT = tile_log2, F = WT - A, NTB = WT - R;
  // tile-index bits as physical positions, ascending; canon(pos) = rank
  int pos[16], np = 0;
  {
    int nf = 0;
    for (int b = 0; b < 64 && np < WT; ++b) {
      bool axis = false;
      for (int a = 0; a < A; ++a) axis = axis || (bt.axbit[a] == b);
      const bool live = (live_mask >> b) & 1ull;
      if (axis) pos[np++] = b;
      else if (live && nf < F) { pos[np++] = b; nf += 1; }
      // (axes above the F-th dense bit are appended when the scan reaches them: np < WT keeps the loop going)
    }
  }
  auto canon = [&](int b) { for (int k = 0; k < np; ++k) if (pos[k] == b) return k; return -1; };
  // slot swizzle: the low three bits (16-byte bank group) are XORed with every higher 3-bit group of the slot index
  auto sw = [](unsigned int x) { return (x & ~7u) | ((x ^ (x >> 3) ^ (x >> 6) ^ (x >> 9)) & 7u); };
  struct Layout { int reg[4]; int thr[MAX_THREAD_BITS]; };
  Layout lay[MAX_LAYOUTS];
  int s_end[MAX_LAYOUTS];
  int nl = 0;
  // the I/O layout: lowest NTB tile-index bits on the threads (lane bits first), the R highest in registers
  Layout io;
  for (int k = 0; k < NTB; ++k) io.thr[k] = pos[k];
  for (int k = 0; k < R; ++k) io.reg[k] = pos[NTB + k];
  // Direct lanes run over the lowest NON-axis bits.  While bit 0, 1 or 2 is one of them a warp access still falls into
  // runs of >= 32 bytes in a few lines (measured: axes {0, 1} at 27 qubits no loss; axes {2, 3, 4} at 33 live qubits 23.5 ms
  // direct against 29.0 ms with the two extra transposes); when bits 0 - 2 are all axes the lanes are >= 128 bytes apart
  // and every access touches 32 lines (config 4 at 33 live qubits: 38 - 50 ms instead of 23 ms per pass).
  int lowest_dense = 64;
  for (int k = 0; k < np; ++k) {
    bool axis = false;
    for (int a = 0; a < A; ++a) axis = axis || (bt.axbit[a] == pos[k]);
    if (!axis) { lowest_dense = pos[k]; break; }
  }
  const bool want_io = (io_mode == 2) || (io_mode == 1 && lowest_dense >= 3);
  auto same_regs = [&](const Layout& x, const Layout& y) {
    for (int k = 0; k < R; ++k) {
      bool in = false;
      for (int m = 0; m < R; ++m) in = in || (x.reg[k] == y.reg[m]);
      if (!in) return false;
    }
    return true;
  };
  auto phase_layout = [&](const Phase& P, bool ascending) {
    Layout L;
    for (int k = 0; k < R; ++k) L.reg[k] = bt.axbit[P.T[k]];
    int rest[16], nr = 0;
    for (int k = 0; k < np; ++k) {
      bool isreg = false;
      for (int m = 0; m < R; ++m) isreg = isreg || (L.reg[m] == pos[k]);
      if (!isreg) rest[nr++] = pos[k];
    }
    // three lane bits with distinct canonical index mod 3 (distinct bank groups after the fold) go first, the rest
    // ascending.  A layout that the tile is loaded / stored in picks them among its five lowest bits: which addresses
    // a warp covers depends on the SET of its five lane bits only, not on their order.
    int nt = 0;
    const int window = ascending ? std::min(nr, 5) : nr;
    bool used_res[3] = {false, false, false};
    bool taken[16] = {false};
    for (int k = 0; k < window && nt < 3; ++k) {
      const int c = canon(rest[k]) % 3;
      if (!used_res[c]) { used_res[c] = true; taken[k] = true; L.thr[nt++] = rest[k]; }
    }
    for (int k = 0; k < nr; ++k) if (!taken[k]) L.thr[nt++] = rest[k];
    return L;
  };
  for (int p = 0; p < nph; ++p) {
    const bool io_first = want_io && p == 0, io_last = want_io && p == nph - 1;
    // without I/O layouts a phase that the tile is loaded / stored in keeps its thread bits in physical order
    Layout L = phase_layout(ph[p], (p == 0 || p == nph - 1) && !want_io);
    const bool match = same_regs(L, io);
    if ((io_first || io_last) && match) for (int k = 0; k < NTB; ++k) L.thr[k] = io.thr[k];   // the phase IS the I/O layout
    if (io_first && !match) { lay[nl] = io; s_end[nl] = 0; nl += 1; }
    lay[nl] = L;
    s_end[nl] = ph[p].s_end;
    nl += 1;
    if (io_last && !match) { lay[nl] = io; s_end[nl] = ph[p].s_end; nl += 1; }
    // stage tables of this phase
    unsigned char regbit[4];
    for (int k = 0; k < R; ++k) regbit[k] = bt.axbit[ph[p].T[k]];
    for (int s = ph[p].s_begin; s < ph[p].s_end; ++s) {
      int la = 0;
      for (int k = 0; k < R; ++k) if (ph[p].T[k] == bt.st[s].axis) la = k;
      out->lax[s] = (unsigned char)la;
      out->st[s].qtab = parity_table(bt.st[s].mq, regbit, R, la);
      out->st[s].vtab = parity_table(bt.st[s].mv, regbit, R, la);
    }
  }
  out->nlayouts = nl;
  out->F = F;
  for (int p = 0; p < MAX_LAYOUTS + 2; ++p) out->s_end[p] = (unsigned char)((p < nl) ? s_end[p] : bt.nstages);
  for (int p = 0; p < nl; ++p) {
    for (int b = 0; b < (1 << R); ++b) {
      uint64_t go = 0;
      unsigned int so = 0;
      for (int k = 0; k < R; ++k)
        if ((b >> k) & 1) { go |= 1ull << lay[p].reg[k]; so |= 1u << canon(lay[p].reg[k]); }
      out->gbits[p][b] = go;
      out->sbits[p][b] = sw(so) * (unsigned int)sizeof(double2);
    }
    if (p == 0) for (int b = 0; b < 16; ++b) out->gload[b] = out->gbits[0][b] * sizeof(double2);
    if (p == nl - 1) for (int b = 0; b < 16; ++b) out->gstore[b] = out->gbits[p][b] * sizeof(double2);
    for (int k = 0; k < MAX_THREAD_BITS; ++k) {
      out->gtid[p][k] = (k < NTB) ? (1ull << lay[p].thr[k]) : 0ull;
      out->stid[p][k] = (k < NTB) ? sw(1u << canon(lay[p].thr[k])) * (unsigned int)sizeof(double2) : 0u;
    }
  }
  return nph;
}

// Plan of the bulk-copy staged pass (k_fused_tma): tile geometry (L low physical bits + the H batch axes above them),
// greedy phases with their three register bits given as tile-index bits, parity tables.  Returns false when the
// layout does not allow it (fewer than 11 live bits, a dead bit inside the contiguous runs, too many phases).
bool plan_tma(const gxsv_t* h, const Batch& bt, TmaBatch* out, Holes* hs) {
  const int n = nlive_bits(h);
  if (n < TILE_LOG2) return false;
  int L = 5;
  for (; L <= TILE_LOG2; ++L) {
    int c = 0;
    for (int a = 0; a < bt.naxes; ++a) c += (bt.axbit[a] >= L) ? 1 : 0;
    if (L + c == TILE_LOG2) break;
  }
  if (L > TILE_LOG2) return false;
  if ((~h->live_mask) & ((1ull << L) - 1ull)) return false;      // a dead bit would break the contiguous runs
  memset(out, 0, sizeof *out);
  out->L = L;
  out->H = TILE_LOG2 - L;
  out->nstages = bt.nstages;
  int hb[8], nh = 0;
  for (int a = 0; a < bt.naxes; ++a) if (bt.axbit[a] >= L) hb[nh++] = bt.axbit[a];
  std::sort(hb, hb + nh);
  for (int i = 0; i < nh; ++i) out->hbit[i] = (unsigned char)hb[i];
  auto tile_bit = [&](int a) {          // tile-index bit of batch axis a
    const int pb = bt.axbit[a];
    if (pb < L) return pb;
    for (int i = 0; i < nh; ++i) if (hb[i] == pb) return L + i;
    return -1;
  };
  auto phys_of = [&](int sbit) { return sbit < L ? sbit : hb[sbit - L]; };
  int nph = 0, i = 0;
  while (i < bt.nstages) {
    if (nph == MAX_PHASES) return false;
    int T[3], nt = 0, j = i;
    for (; j < bt.nstages; ++j) {
      const int sbit = tile_bit(bt.st[j].axis);
      bool in = false;
      for (int k = 0; k < nt; ++k) in = in || (T[k] == sbit);
      if (in) continue;
      if (nt == 3) break;
      T[nt++] = sbit;
    }
    for (int sbit = TILE_LOG2 - 1; sbit >= 0 && nt < 3; --sbit) {   // pad with the highest free tile-index bits
      bool in = false;
      for (int k = 0; k < nt; ++k) in = in || (T[k] == sbit);
      if (!in) T[nt++] = sbit;
    }
    std::sort(T, T + 3);
    unsigned char regbit[3];
    for (int k = 0; k < 3; ++k) { out->sb[nph][k] = (unsigned char)T[k]; regbit[k] = (unsigned char)phys_of(T[k]); }
    out->s_begin[nph] = (unsigned char)i;
    out->s_end[nph] = (unsigned char)j;
    for (int s2 = i; s2 < j; ++s2) {
      out->st[s2] = bt.st[s2];
      out->nscale[s2] = bt.nscale[s2];
      const int sbit = tile_bit(bt.st[s2].axis);
      int la = 0;
      for (int k = 0; k < 3; ++k) if (T[k] == sbit) la = k;
      out->lax[s2] = (unsigned char)la;
      out->st[s2].qtab = parity_table(bt.st[s2].mq, regbit, 3, la);
      out->st[s2].vtab = parity_table(bt.st[s2].mv, regbit, 3, la);
    }
    nph += 1;
    i = j;
  }
  out->nphases = nph;
  // tiles are enumerated over the live bits at / above L that are not batch axes
  hs->n = 0;
  for (int b = L; b < h->phys_bits; ++b) {
    bool skip = !((h->live_mask >> b) & 1ull);
    for (int k = 0; k < nh; ++k) skip = skip || (hb[k] == b);
    if (skip) hs->pos[hs->n++] = (unsigned char)b;
  }
  return true;
}

// Run the queued projection stages (k_fused_batch / k_fused_tile).  The layout bookkeeping was done when they were queued.
int flush_batch(gxsv_t* h) {
  if (h->batch.nstages == 0) return GXSV_OK;
  Batch bt = h->batch;
  h->batch.nstages = 0;
  h->batch.naxes = 0;
  bt.st[bt.nstages - 1].acc = 1;             // the last norm of a pass is always measured: it sets the normalisation
  h->open_notes[bt.nstages - 1].an = 0.0;
  Holes hs;
  hs.n = 0;
  for (int b = 0; b < h->phys_bits; ++b) {
    bool skip = !((h->live_mask >> b) & 1ull);
    for (int a = 0; a < bt.naxes; ++a) skip = skip || (bt.axbit[a] == b);
    if (skip) hs.pos[hs.n++] = (unsigned char)b;
  }
  const int n = __builtin_popcountll(h->live_mask);
  const uint64_t ngroups = 1ull << (n - bt.naxes);
  unsigned long long seq;
  Res* slot = next_slot(h, &seq);
  const int grid = grid_for(h, ngroups, 1);
  const int fm = h->dist ? FIN_ONESHOT : FIN_BATCH;
  if (bt.nstages == 1) {
    // a batch of one (e.g. every projection is preceded by a probability query): the plain butterfly kernel is leaner
    const Stage& st = bt.st[0];
    const uint64_t npair = 1ull << (n - 1);
    // the host scalar already holds the factor f taken out of a kind 0 / 1 stage: hand k_fused the matrix without it
    double2 m00 = st.m00, m01 = st.m01, m10 = st.m10, m11 = st.m11;
    if (st.kind != 2) {
      const double rho = st.neg ? -1.0 : 1.0;
      const double2 one = make_double2(1.0, 0.0), ra = make_double2(rho * st.a.x, rho * st.a.y);
      if (st.kind == 0) { m00 = one; m01 = st.a; m10 = make_double2(rho, 0.0); m11 = make_double2(-ra.x, -ra.y); }
      else { m00 = st.a; m01 = one; m10 = ra; m11 = make_double2(-rho, 0.0); }
    }
    Launch L(h, GXSV_K_FUSED, 2.0 * 16.0 * std::ldexp(1.0, n), /*frame_aware=*/true);   // a pending frame was issued AFTER these stages
    k_fused<U_DEF><<<grid_for(h, npair, U_DEF), TPB, 0, h->stream>>>(h->psi, hs, bt.axbit[0], m00, m01, m10, m11, st.mq,
                                                                    st.mv, npair, h->partials, h->ctrl, slot, seq,
                                                                    h->dist ? FIN_ONESHOT : FIN_CONTRACT, bt.nscale[0]);
  } else if (h->use_tma && plan_tma(h, bt, &g_tma, &g_tma_holes)) {
    const uint64_t ntiles = 1ull << (n - TILE_LOG2);
    const size_t dyn = (sizeof(double2) << TILE_LOG2) * TMA_BUFS;
    if (!h->attr_tma) {      // per handle: function attributes belong to the device context
      CU(cudaFuncSetAttribute(k_fused_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      h->attr_tma = true;
    }
    const uint64_t tgrid = std::min<uint64_t>(ntiles, (uint64_t)h->sm_count * 2);   // persistent: 2 resident CTAs per SM
    Launch L(h, GXSV_K_FUSED, 2.0 * 16.0 * std::ldexp(1.0, n), /*frame_aware=*/true);   // a pending frame was issued AFTER these stages
    k_fused_tma<<<(int)tgrid, TPB, dyn, h->stream>>>(h->psi, g_tma_holes, g_tma, ntiles, h->partials, h->ctrl, slot, seq, fm);
  } else if (bt.naxes <= 3 || bt.naxes > h->tile_regs + 3 || n < 8 + h->tile_regs + 1) {
    if (bt.naxes > 3) return fail(h, GXSV_EVALUE, "internal: batch on %d axes cannot run", bt.naxes);
    for (int k = 0; k < bt.nstages; ++k) {
      bt.st[k].qtab = parity_table(bt.st[k].mq, bt.axbit, bt.naxes, bt.st[k].axis);
      bt.st[k].vtab = parity_table(bt.st[k].mv, bt.axbit, bt.naxes, bt.st[k].axis);
    }
    Launch L(h, GXSV_K_FUSED, 2.0 * 16.0 * std::ldexp(1.0, n), /*frame_aware=*/true);   // a pending frame was issued AFTER these stages
    // chunked CTA -> index mapping only when every CTA still gets >= 8 chunks (else plain grid stride)
    int gl = h->chunk_log2;
    while (gl > 0 && (ngroups >> (8 + gl)) < (uint64_t)grid * 8) --gl;
    if (bt.naxes == 1) k_fused_batch<1><<<grid, TPB, 0, h->stream>>>(h->psi, hs, bt, ngroups, gl, h->partials, h->ctrl, slot, seq, fm);
    else if (bt.naxes == 2) k_fused_batch<2><<<grid, TPB, 0, h->stream>>>(h->psi, hs, bt, ngroups, gl, h->partials, h->ctrl, slot, seq, fm);
    else k_fused_batch<3><<<grid, TPB, 0, h->stream>>>(h->psi, hs, bt, ngroups, gl, h->partials, h->ctrl, slot, seq, fm);
  } else {
    static thread_local TileBatch tb;
    memset(&tb, 0, sizeof tb);
    tb.nstages = bt.nstages;
    tb.naxes = bt.naxes;
    for (int k = 0; k < bt.nstages; ++k) { tb.st[k] = bt.st[k]; tb.nscale[k] = bt.nscale[k]; }
    const int R = h->tile_regs;
    // warps per tile: private warp tiles while the runs stay >= 256 bytes (up to 5 axes at R = 4), else four warps share
    // a tile (runs of 2^(7 + R - A) amplitudes); GXSV_TILE_WARPS / GXSV_OPT_TILE_WARPS override (1, 4 or 8)
    int TW = h->tile_warps;
    if (TW == 0) TW = (bt.naxes >= R + 2) ? 4 : 1;
    if (R == 3) TW = 1;
    const int ltw = (TW == 8) ? 3 : (TW == 4) ? 2 : 0;
    const int nphases = plan_phases(bt, &tb, R, h->live_mask, 5 + R + ltw, h->tile_io);
    if (nphases > MAX_PHASES) return fail(h, GXSV_EVALUE, "internal: batch needs %d phases", nphases);
    for (int k = 0; k < bt.nstages; ++k) {      // dispatch codes and packed operands of the stages (stage_by_variant)
      const Stage& st = tb.st[k];
      unsigned int variant;
      if (st.kind == 0 && st.mv == 0ull && !st.neg) variant = st.acc ? SV_FAST_ACC : (st.qtab == 0ull ? SV_UNIQ : SV_FAST);
      else variant = (st.kind == 0) ? SV_K0_SIGNED : (st.kind == 1) ? SV_K1 : SV_GENERAL;
      tb.code[k] = (unsigned char)(32u * tb.lax[k] + (st.acc ? 16u : 0u) + variant);
      tb.hot[k].a = st.a;
      tb.hot[k].qtab = st.qtab;
      tb.hot[k].mq = st.mq;
    }
    const uint64_t ntiles = 1ull << (n - (8 + R));            // CTA iterations: 8 warp tiles of 2^(5+R) amplitudes
    const size_t dyn = sizeof(double2) << (8 + R);
    if (!h->attr_tile) {     // per handle: function attributes belong to the device context
      const int dyn11 = (int)(sizeof(double2) << 11), dyn12 = (int)(sizeof(double2) << 12);
      CU(cudaFuncSetAttribute(k_fused_tile<3, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn11));
      CU(cudaFuncSetAttribute(k_fused_tile<4, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn12));
      CU(cudaFuncSetAttribute(k_fused_tile<4, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn12));
      CU(cudaFuncSetAttribute(k_fused_tile<4, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn12));
      CU(cudaFuncSetAttribute(k_fused_tile<4, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn12));
      CU(cudaFuncSetAttribute(k_fused_tile<4, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn12));
      h->attr_tile = true;
    }
    uint64_t tgrid = std::min<uint64_t>(ntiles, (uint64_t)h->sm_count * (R == 3 ? 4 : 2));   // persistent: resident CTAs only
    int gl = h->chunk_log2;
    while (gl > 0 && (ntiles >> gl) < tgrid * 8) --gl;
    // GXSV_BATCH_LOG=1: one line per pass on stderr (shape, axis bits); =2: also the pass time (synchronises every pass)
    static const int log_shapes = std::getenv("GXSV_BATCH_LOG") ? std::atoi(std::getenv("GXSV_BATCH_LOG")) : 0;
    cudaEvent_t le0 = nullptr, le1 = nullptr;
    if (log_shapes >= 2) { cudaEventCreate(&le0); cudaEventCreate(&le1); cudaEventRecord(le0, h->stream); }
    Launch L(h, GXSV_K_FUSED, 2.0 * 16.0 * std::ldexp(1.0, n), /*frame_aware=*/true);   // a pending frame was issued AFTER these stages
#define GXSV_TILE_LAUNCH(RR, WW, SS) \
  k_fused_tile<RR, WW, SS><<<(int)tgrid, TPB, dyn, h->stream>>>(h->psi, hs, tb, ntiles, gl, h->partials, h->ctrl, slot, seq, fm)
    const bool sp = h->stage_prefetch != 0;
    if (R == 3) GXSV_TILE_LAUNCH(3, 1, false);
    else if (TW == 8) GXSV_TILE_LAUNCH(4, 8, true);
    else if (TW == 4 && sp) GXSV_TILE_LAUNCH(4, 4, true);
    else if (TW == 4) GXSV_TILE_LAUNCH(4, 4, false);
    else if (sp) GXSV_TILE_LAUNCH(4, 1, true);
    else GXSV_TILE_LAUNCH(4, 1, false);
#undef GXSV_TILE_LAUNCH
    if (log_shapes) {
      float ms = 0.f;
      if (log_shapes >= 2) {
        cudaEventRecord(le1, h->stream);
        cudaEventSynchronize(le1);
        cudaEventElapsedTime(&ms, le0, le1);
        cudaEventDestroy(le0);
        cudaEventDestroy(le1);
      }
      int nuni = 0;
      for (int k = 0; k < bt.nstages; ++k) nuni += ((tb.code[k] & 7u) == SV_UNIQ) ? 1 : 0;
      fprintf(stderr, "gxsv-pass n=%d axes=%d stages=%d phases=%d layouts=%d uniform=%d tw=%d ms=%.4f bits=", n, bt.naxes,
              bt.nstages, nphases, tb.nlayouts, nuni, TW, ms);
      for (int a = 0; a < bt.naxes; ++a) fprintf(stderr, "%d%s", (int)bt.axbit[a], a + 1 < bt.naxes ? "," : "\n");
    }
  }
  Deferred d;
  d.first = seq;
  d.second = bt.nstages;
  for (int k = 0; k < bt.nstages; ++k) d.note[k] = h->open_notes[k];
  h->deferred.push_back(d);
  return check_launch(h, "k_fused_batch");
}

// Non-strict mode: verify the norms of the projections issued so far (statevec.py:1192-1193, raised late).
int check_deferred(gxsv_t* h) {
  int rc = flush_batch(h);
  if (rc) return rc;
  if (h->deferred.empty() || h->dist) return GXSV_OK;   // sharded: the caller collects the norms (gxsv_take_norms)
  std::vector<Deferred> todo;
  todo.swap(h->deferred);
  for (auto& d : todo) {
    double v[RES_VALUES];
    rc = wait_slot(h, d.first, v);
    if (rc) return rc;
    double prev = 1.0;
    for (int k = 0; k < d.second; ++k) {
      const StageNote& nt = d.note[k];
      if (nt.an > 0.0) v[k] = prev * nt.an;      // not accumulated: known analytically
      const double p = v[k] / prev;   // norm^2 of the kept (larger) row for a normalised input
      const double n0 = nt.big1 ? p * nt.ratio : p, n1 = nt.big1 ? p : p * nt.ratio;
      if (!(n0 > 1e-10) && !(n1 > 1e-10))
        return fail(h, GXSV_EZERONORM, "Attempted to remove a qubit from 0-norm statevector (detected at sync).");
      if (nt.big1) {
        const bool actual = n0 > 1e-10;   // the reference would have kept half 0
        if (actual != nt.assumed) h->hscale *= actual ? nt.phi : std::conj(nt.phi);
      }
      prev = v[k];
    }
  }
  return GXSV_OK;
}

// Make the pending Pauli frame real: one swap pass for all X flips, one sign pass for all Z's (X first: the frame is
// kept in the order Z^z X^x per qubit).
int settle_frame(gxsv_t* h) {
  if (!(h->fx | h->fz)) return GXSV_OK;
  const uint64_t fx = h->fx, fz = h->fz;
  h->fx = h->fz = 0;                         // cleared first: the launches below go through Launch
  const int n = nlive_bits(h);
  const double S = 16.0 * std::ldexp(1.0, n);
  if (fx && n >= 1) {
    const int b0 = __builtin_ctzll(fx);
    Holes hs = make_holes(h, b0);
    const uint64_t npair = 1ull << (n - 1);
    Launch L(h, GXSV_K_APPLY1Q, 2.0 * S);
    k_xflip<U_DEF><<<grid_for(h, npair, U_DEF), TPB, 0, h->stream>>>(h->psi, hs, fx, npair);
  }
  if (fz) {
    Holes hs = make_holes(h);
    static thread_local EdgeList el;
    el.n = 0;
    const uint64_t nlive = 1ull << n;
    Launch L(h, GXSV_K_APPLY1Q, S);
    k_diag_edges<U_DEF><<<grid_for(h, nlive, U_DEF), TPB, 0, h->stream>>>(h->psi, hs, el, fz, nlive);
  }
  return check_launch(h, "settle_frame");
}

void trim_extent(gxsv_t* h) {
  while (h->phys_bits > 0 && !((h->live_mask >> (h->phys_bits - 1)) & 1ull)) h->phys_bits -= 1;
}

int check_q(gxsv_t* h, int q) {
  if (q < 0 || q >= (int)h->q.size())
    return fail(h, GXSV_EINDEX, "Qubit index %d out of range [0, %d]", q, (int)h->q.size() - 1);
  return GXSV_OK;
}

double state_bytes(const gxsv_t* h) { return 16.0 * std::ldexp(1.0, nlive_bits(h)); }

// --- queued CZs --------------------------------------------------------------
Qubit* by_id(gxsv_t* h, int id) {
  for (auto& qb : h->q) if (qb.id == id) return &qb;
  return nullptr;
}

// CZ^2 = 1: canonical order, cancel duplicates
void dedupe_pending(gxsv_t* h) {
  std::vector<std::pair<int, int>> edges;
  for (auto e : h->pend) {
    if (e.first > e.second) std::swap(e.first, e.second);
    auto it = std::find(edges.begin(), edges.end(), e);
    if (it != edges.end()) edges.erase(it); else edges.push_back(e);
  }
  h->pend.swap(edges);
}

// Issue the queued CZs whose two qubits are materialised (all of them, or only those touching
// qubit `touching` when >= 0) as diagonal phase passes: stars sharing a pivot go through
// k_cz_star (touches a quarter of the amplitudes), anything else through one general pass.
int flush_real_edges(gxsv_t* h, int touching = -1) {
  if (h->pend.empty()) return GXSV_OK;
  dedupe_pending(h);
  std::vector<std::pair<int, int>> edges, keep;
  for (auto& e : h->pend) {
    Qubit* a = by_id(h, e.first);
    Qubit* b = by_id(h, e.second);
    const bool real = a && b && a->phys >= 0 && b->phys >= 0;
    if (real && (touching < 0 || e.first == touching || e.second == touching)) {
      int pa = a->phys, pb = b->phys;
      if (pa > pb) std::swap(pa, pb);
      edges.push_back({pa, pb});
    } else {
      keep.push_back(e);
    }
  }
  h->pend.swap(keep);
  const int n = nlive_bits(h);
  const double S = state_bytes(h);
  while (!edges.empty()) {
    // vertex of highest degree becomes the pivot of a star
    int deg[64] = {0};
    for (auto& e : edges) { deg[e.first]++; deg[e.second]++; }
    int m = 0;
    for (int b = 1; b < 64; ++b) if (deg[b] > deg[m]) m = b;
    const bool single_star = (deg[m] == (int)edges.size());
    if (!single_star && edges.size() <= 64 && edges.size() > 2) {
      // several stars: one general diagonal pass over the whole state is cheaper than >= 3 star passes
      EdgeList el;
      el.n = (int)edges.size();
      for (int i = 0; i < el.n; ++i) { el.a[i] = (unsigned char)edges[i].first; el.b[i] = (unsigned char)edges[i].second; }
      const uint64_t nlive = 1ull << n;
      Holes hs = make_holes(h);
      {
        Launch L(h, GXSV_K_CZ, 1.5 * S);
        k_diag_edges<U_DEF><<<grid_for(h, nlive, U_DEF), TPB, 0, h->stream>>>(h->psi, hs, el, 0ull, nlive);
      }
      edges.clear();
      return check_launch(h, "k_diag_edges");
    }
    uint64_t nbr = 0;
    std::vector<std::pair<int, int>> rest;
    for (auto& e : edges) {
      if (e.first == m) nbr |= 1ull << e.second;
      else if (e.second == m) nbr |= 1ull << e.first;
      else rest.push_back(e);
    }
    edges.swap(rest);
    const int hb = 63 - __builtin_clzll(nbr);
    const uint64_t nbr_rest = nbr & ~(1ull << hb);
    const uint64_t nquarter = (n >= 2) ? (1ull << (n - 2)) : 0;
    if (nquarter == 0) continue;
    Holes hs = make_holes(h, m, hb);
    // algorithmic bytes: the flipped quarter read + written; a participating bit below
    // the 32-byte sector (bit 0) forces whole sectors of the pivot half to move
    int lowbit = std::min(m, __builtin_ctzll(nbr));
    const double bytes = (lowbit == 0) ? S : 0.5 * S;
    {
      Launch L(h, GXSV_K_CZ, bytes);
      k_cz_star<U_DEF><<<grid_for(h, nquarter, U_DEF), TPB, 0, h->stream>>>(h->psi, hs, m, hb, nbr_rest, nquarter);
    }
    int rc = check_launch(h, "k_cz_star");
    if (rc) return rc;
  }
  return GXSV_OK;
}

// choose the physical bit for a new qubit: lowest hole, else the next bit on top
int alloc_bit(gxsv_t* h, int* bit) {
  for (int b = 0; b < h->phys_bits; ++b)
    if (!((h->live_mask >> b) & 1ull)) { *bit = b; return GXSV_OK; }
  if (h->phys_bits + 1 > h->cap_bits) {
    int rc = gxsv_ensure_capacity(h, h->phys_bits + 1);
    if (rc) return rc;
  }
  *bit = h->phys_bits;
  return GXSV_OK;
}

// Give the virtual qubit at logical index qi a physical bit (k_fill): psi <- psi (x) (s0, s1).
int materialize(gxsv_t* h, int qi) {
  Qubit& qb = h->q[qi];
  if (qb.phys >= 0) return GXSV_OK;
  const cplx s0 = qb.s0, s1 = qb.s1;
  int hb;
  int rc = alloc_bit(h, &hb);
  if (rc) return rc;
  const int n = nlive_bits(h);
  const uint64_t nlive = 1ull << n;
  const double S = 16.0 * (double)nlive;
  Holes hs = make_holes(h, hb);
  const int grid = grid_for(h, nlive, U_DEF);
  const double a0 = std::abs(s0), a1 = std::abs(s1);
  if (a0 >= a1 || a0 > 0.25) {
    // defer s0 to the host scalar; upper half = (s1/s0) * lower half
    const cplx f1 = s1 / s0;
    Launch L(h, GXSV_K_FILL, 2.0 * S);
    if (f1 == cplx(1.0, 0.0))
      k_fill<U_DEF, 0><<<grid, TPB, 0, h->stream>>>(h->psi, hs, hb, d2(1.0), d2(1.0), nlive);
    else
      k_fill<U_DEF, 1><<<grid, TPB, 0, h->stream>>>(h->psi, hs, hb, d2(1.0), d2(f1), nlive);
    h->hscale *= s0;
  } else if (h->dist) {
    // sharded: s1 may carry a rank-dependent sign, so nothing of it may go to the (rank-uniform) host scalar
    Launch L(h, GXSV_K_FILL, 3.0 * S);
    k_fill<U_DEF, 2><<<grid, TPB, 0, h->stream>>>(h->psi, hs, hb, d2(s0), d2(s1), nlive);
  } else {
    // |s0| small (e.g. |1>): defer s1 instead, lower half scaled by s0/s1
    const cplx f0 = s0 / s1;
    Launch L(h, GXSV_K_FILL, 3.0 * S);
    k_fill<U_DEF, 2><<<grid, TPB, 0, h->stream>>>(h->psi, hs, hb, d2(f0), d2(1.0), nlive);
    h->hscale *= s1;
  }
  rc = check_launch(h, "k_fill");
  if (rc) return rc;
  h->live_mask |= 1ull << hb;
  if (hb >= h->phys_bits) h->phys_bits = hb + 1;
  h->q[qi].phys = hb;
  return GXSV_OK;
}

// everything queued goes to the device: all virtual qubits, then all CZs
int flush_pending(gxsv_t* h) {
  for (size_t i = 0; i < h->q.size(); ++i) {
    int rc = materialize(h, (int)i);
    if (rc) return rc;
  }
  return flush_real_edges(h, -1);
}

bool has_edge(const gxsv_t* h, int id) {
  for (auto& e : h->pend) if (e.first == id || e.second == id) return true;
  return false;
}

// Make qubit qi ready for an arbitrary (non-diagonal) operation: materialise it and its virtual
// CZ partners, then issue the CZs that touch it.  Queued work elsewhere stays queued.
int settle(gxsv_t* h, int qi) {
  if (h->fusion < 2) return flush_pending(h);
  dedupe_pending(h);
  const int id = h->q[qi].id;
  int rc = materialize(h, qi);
  if (rc) return rc;
  for (auto& e : h->pend) {
    if (e.first != id && e.second != id) continue;
    const int other = (e.first == id) ? e.second : e.first;
    for (size_t i = 0; i < h->q.size(); ++i)
      if (h->q[i].id == other) { rc = materialize(h, (int)i); if (rc) return rc; }
  }
  return flush_real_edges(h, id);
}

int add_one(gxsv_t* h, cplx s0, cplx s1) {
  h->q.push_back({h->next_id++, -1, s0, s1});
  if (h->fusion >= 2) return GXSV_OK;  // lazy: stays virtual until something needs it
  int rc = flush_real_edges(h, -1);
  if (rc) return rc;
  return materialize(h, (int)h->q.size() - 1);
}

// Launch the contraction of physical bit p with bra (c0, c1) (host scalar already folded in by
// the caller) and update the layout.  `qidx` = logical index of the measured qubit.
int launch_contract(gxsv_t* h, int qidx, cplx c0, cplx c1, uint64_t mq, unsigned long long* seq_out) {
  const int p = h->q[qidx].phys;
  const int n = nlive_bits(h);
  const uint64_t nout = 1ull << (n - 1);
  const double S = state_bytes(h);
  std::vector<int> L;  // live physical bits, ascending
  for (int b = 0; b < h->phys_bits; ++b) if ((h->live_mask >> b) & 1ull) L.push_back(b);
  const bool tile = (n <= TILE_BITS) || (p < HOLE_MIN);
  unsigned long long seq;
  Res* slot = next_slot(h, &seq);
  *seq_out = seq;
  Holes hs = make_holes(h, p);
  if (!tile) {
    const int grid = grid_for(h, nout, U_DEF);
    {
      Launch Lc(h, GXSV_K_CONTRACT, 1.5 * S);
      k_contract<U_DEF, false><<<grid, TPB, 0, h->stream>>>(h->psi, hs, hs, p, d2(c0), d2(c1), mq, nout,
                                                           h->partials, h->ctrl, slot, seq,
                                                           (h->dist ? (h->strict ? FIN_CONTRACT_DIST : FIN_ONESHOT) : FIN_CONTRACT));
    }
    h->live_mask &= ~(1ull << p);
  } else {
    const int K = std::min(TILE_BITS, n);
    const int top = L[K - 1];  // the tile's top live bit becomes the hole
    int t = 0;
    while (L[t] != p) ++t;
    const uint64_t live_after = h->live_mask & ~(1ull << top);
    Holes hd = holes_of_mask(live_after, h->phys_bits);
    // every CTA must own whole tiles: grid = number of 1024-output chunks (capped, grid-stride keeps alignment)
    const int grid = grid_for(h, nout, U_DEF);
    {
      Launch Lc(h, GXSV_K_CONTRACT, 1.5 * S);
      k_contract<U_DEF, true><<<grid, TPB, 0, h->stream>>>(h->psi, hs, hd, p, d2(c0), d2(c1), mq, nout,
                                                          h->partials, h->ctrl, slot, seq,
                                                          (h->dist ? (h->strict ? FIN_CONTRACT_DIST : FIN_ONESHOT) : FIN_CONTRACT));
    }
    // qubits on L[t+1 .. K-1] slide down one live slot
    for (auto& qb : h->q) {
      if (qb.phys < 0) continue;
      for (int i = t + 1; i <= K - 1; ++i)
        if (qb.phys == L[i]) { qb.phys = L[i - 1]; break; }
    }
    h->live_mask = live_after;
  }
  int rc = check_launch(h, "k_contract");
  if (rc) return rc;
  h->q.erase(h->q.begin() + qidx);
  trim_extent(h);
  return GXSV_OK;
}

int ensure_stage(gxsv_t* h, size_t amps) {
  if (h->stage_amps >= amps) return GXSV_OK;
  for (int i = 0; i < 2; ++i) if (h->stage[i]) { cudaFree(h->stage[i]); h->stage[i] = nullptr; }
  h->stage_amps = 0;
  for (int i = 0; i < 2; ++i) CU(cudaMalloc(&h->stage[i], amps * sizeof(double2)));
  h->stage_amps = amps;
  return GXSV_OK;
}

int upload_tables(gxsv_t* h) {
  // canonical index bit (n-1-q) of logical qubit q  ->  physical bit q.phys
  const int n = (int)h->q.size();
  static thread_local GatherTables t;
  for (int byte = 0; byte < 5; ++byte) {
    for (int v = 0; v < 256; ++v) {
      uint64_t p = 0;
      for (int b = 0; b < 8; ++b) {
        if (!((v >> b) & 1)) continue;
        const int bit = byte * 8 + b;  // canonical bit index
        if (bit >= n) continue;
        const int ql = n - 1 - bit;
        p |= 1ull << h->q[ql].phys;
      }
      t.t[byte][v] = p;
    }
  }
  CU(cudaMemcpyAsync(h->tabs_dev, &t, sizeof t, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return GXSV_OK;
}

int init_common(gxsv_t* h) {
  CU(cudaSetDevice(h->device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, h->device));
  h->sm_count = prop.multiProcessorCount;
  CU(cudaMalloc(&h->ctrl, sizeof(Ctrl)));
  CU(cudaMalloc(&h->partials, sizeof(double) * RES_VALUES * (size_t)(h->sm_count * 8 + 8)));
  CU(cudaMalloc(&h->tabs_dev, sizeof(GatherTables)));
  CU(cudaMalloc(&h->op_dev, sizeof(double2) * 256));
  CU(cudaHostAlloc(&h->ring_host, sizeof(Res) * RING, cudaHostAllocMapped));
  memset(h->ring_host, 0, sizeof(Res) * RING);
  CU(cudaHostGetDevicePointer((void**)&h->ring_dev, h->ring_host, 0));
  memset(&h->stats, 0, sizeof h->stats);
  memset(&h->batch, 0, sizeof h->batch);
  return gxsv_reset(h);
}

// Phase convention of a non-strict projection: guess which half the reference would keep, apply the phase now,
// remember enough to correct it once the norm is known (check_deferred).
StageNote phase_note(gxsv_t* h, int big, double r0, double r1, cplx m00, cplx m01, cplx m10, cplx m11) {
  StageNote nt;
  nt.big1 = (big == 1);
  if (!nt.big1) { nt.ratio = r0 > 0.0 ? r1 / r0 : 0.0; return nt; }
  nt.ratio = r0 / r1;
  const cplx ub = std::abs(m10) >= std::abs(m11) ? m10 : m11;
  const cplx up = std::abs(m10) >= std::abs(m11) ? m00 : m01;
  const cplx ratio = up / ub;
  nt.phi = std::abs(ratio) > 0.0 ? ratio / std::abs(ratio) : cplx(1.0);
  nt.assumed = r0 > 1e-10 * r1;
  if (nt.assumed && !h->dist) h->hscale *= nt.phi;   // sharded: the orchestrator owns the convention
  return nt;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

int gxsv_version(void) { return GXSV_VERSION; }

const char* gxsv_last_error(gxsv_t* h) { return h ? h->err.c_str() : g_err.c_str(); }

int gxsv_create(int max_qubits, int device, void* stream, gxsv_t** out) {
  gxsv_t* h = nullptr;
  if (!out) return fail(h, GXSV_EVALUE, "out is NULL");
  if (max_qubits < 0 || max_qubits > 40) return fail(h, GXSV_EVALUE, "`max_qubits` must be in [0, 40], got %d", max_qubits);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(h, GXSV_ECUDA, "no CUDA device available (%s): the gxsv engine has no CPU fallback",
                cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(h, GXSV_EVALUE, "device %d out of range [0, %d)", device, ndev);
  h = new gxsv();
  h->device = device;
  if (const char* ax = std::getenv("GXSV_BATCH_AXES")) {   // tuning knob for the batched pass (3 or 4), see kernels.cuh
    const int v = std::atoi(ax);
    if (v >= 1 && v <= MAX_TILE_AXES) h->batch_axes = v;
  }
  if (const char* tm = std::getenv("GXSV_TMA")) h->use_tma = std::atoi(tm) ? 1 : 0;
  if (const char* mp = std::getenv("GXSV_MAX_PHASES")) { const int v = std::atoi(mp); if (v >= 1 && v <= MAX_PHASES) h->max_phases = v; }
  if (const char* lf = std::getenv("GXSV_LAZY_FRAME")) h->lazy_frame = std::atoi(lf) != 0;
  if (const char* sp = std::getenv("GXSV_STAGE_PREFETCH")) h->stage_prefetch = std::atoi(sp) != 0;
  if (const char* tw = std::getenv("GXSV_TILE_WARPS")) { const int v = std::atoi(tw); if (v == 0 || v == 1 || v == 4 || v == 8) h->tile_warps = v; }
  if (const char* tp = std::getenv("GXSV_TILE_IO")) { const int v = std::atoi(tp); if (v >= 0 && v <= 2) h->tile_io = v; }
  if (const char* dv = std::getenv("GXSV_DEFER_VCZ")) h->defer_partner_cz = std::atoi(dv) != 0;
  if (const char* tr = std::getenv("GXSV_TILE_REGS")) h->tile_regs = (std::atoi(tr) == 3) ? 3 : 4;
  if (const char* an = std::getenv("GXSV_ANALYTIC_NORMS")) h->analytic_norms = std::atoi(an) ? 1 : 0;
  if (const char* cl = std::getenv("GXSV_CHUNK_LOG2")) {
    const int v = std::atoi(cl);
    if (v >= 0 && v <= 10) h->chunk_log2 = v;
  }
  h->stream = (cudaStream_t)stream;
  h->cap_bits = max_qubits;
  h->own_buf = true;
  cudaSetDevice(device);
  e = cudaMalloc(&h->psi, sizeof(double2) << max_qubits);
  if (e != cudaSuccess) {
    cudaGetLastError();  // a failed allocation is not sticky, but it stays in the last-error slot until read
    int rc = fail(nullptr, GXSV_ENOMEM, "cudaMalloc of %.3f GiB failed: %s", std::ldexp(16.0, max_qubits - 30),
                  cudaGetErrorString(e));
    delete h;
    return rc;
  }
  int rc = init_common(h);
  if (rc) { g_err = h->err; gxsv_destroy(h); return rc; }
  *out = h;
  return GXSV_OK;
}

int gxsv_create_external(void* device_buffer, int cap_qubits, int device, void* stream, gxsv_t** out) {
  gxsv_t* h = nullptr;
  if (!out || !device_buffer) return fail(h, GXSV_EVALUE, "NULL argument");
  if (cap_qubits < 0 || cap_qubits > 40) return fail(h, GXSV_EVALUE, "`cap_qubits` must be in [0, 40], got %d", cap_qubits);
  {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      return fail(h, GXSV_ECUDA, "no CUDA device available (%s): the gxsv engine has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(h, GXSV_EVALUE, "device %d out of range [0, %d)", device, ndev);
  }
  h = new gxsv();
  h->device = device;
  h->stream = (cudaStream_t)stream;
  h->cap_bits = cap_qubits;
  h->own_buf = false;
  h->psi = (double2*)device_buffer;
  int rc = init_common(h);
  if (rc) { g_err = h->err; gxsv_destroy(h); return rc; }
  *out = h;
  return GXSV_OK;
}

int gxsv_destroy(gxsv_t* h) {
  if (!h) return GXSV_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  prof_collect(h);
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  if (h->own_buf && h->psi) cudaFree(h->psi);
  if (h->ctrl) cudaFree(h->ctrl);
  if (h->partials) cudaFree(h->partials);
  if (h->tabs_dev) cudaFree(h->tabs_dev);
  if (h->op_dev) cudaFree(h->op_dev);
  if (h->ring_host) cudaFreeHost(h->ring_host);
  for (int i = 0; i < 2; ++i) if (h->stage[i]) cudaFree(h->stage[i]);
  for (int i = 0; i < 8; ++i) if (h->peer_psi[i]) cudaIpcCloseMemHandle(h->peer_psi[i]);
  delete h;
  return GXSV_OK;
}

int gxsv_reset(gxsv_t* h) {
  CU(cudaSetDevice(h->device));
  {
    // projections still in flight are verified before the register is thrown away
    int rc = (h->ctrl && h->ring_host) ? check_deferred(h) : GXSV_OK;
    h->batch.nstages = 0;
    h->batch.naxes = 0;
    h->deferred.clear();
    if (rc) return rc;
  }
  h->q.clear();
  h->pend.clear();
  h->next_id = 0;
  h->phys_bits = 0;
  h->live_mask = 0;
  h->hscale = 1.0;
  h->fx = h->fz = 0;
  const double2 one = make_double2(1.0, 0.0);
  const Ctrl c0 = {1.0, 0u, 0u};
  CU(cudaMemcpyAsync(h->psi, &one, sizeof one, cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->ctrl, &c0, sizeof c0, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return GXSV_OK;
}

int gxsv_set_option(gxsv_t* h, int key, int value) {
  switch (key) {
    case GXSV_OPT_FUSION:
      if (value < 0 || value > 2) return fail(h, GXSV_EVALUE, "fusion must be 0, 1 or 2");
      { int rc = gxsv_flush(h); if (rc) return rc; }
      h->fusion = value;
      return GXSV_OK;
    case GXSV_OPT_STRICT:
      if (value) { int rc = check_deferred(h); if (rc) return rc; }
      h->strict = value ? 1 : 0;
      return GXSV_OK;
    case GXSV_OPT_PROFILE:
      if (!value) prof_collect(h);
      h->profile = value ? 1 : 0;
      return GXSV_OK;
    case GXSV_OPT_DIST: h->dist = value ? 1 : 0; return GXSV_OK;
    case GXSV_OPT_BATCH:
      if (value < 0 || value > MAX_STAGES) return fail(h, GXSV_EVALUE, "batch must be in [0, %d]", (int)MAX_STAGES);
      { int rc = flush_batch(h); if (rc) return rc; }
      h->batch_max = value;
      return GXSV_OK;
    case GXSV_OPT_BATCH_AXES:
      if (value < 1 || value > MAX_TILE_AXES) return fail(h, GXSV_EVALUE, "batch axes must be in [1, %d]", (int)MAX_TILE_AXES);
      { int rc = flush_batch(h); if (rc) return rc; }
      h->batch_axes = value;
      return GXSV_OK;
    case GXSV_OPT_CHUNK_LOG2:
      if (value < 0 || value > 10) return fail(h, GXSV_EVALUE, "chunk log2 must be in [0, 10]");
      h->chunk_log2 = value;
      return GXSV_OK;
    case GXSV_OPT_TMA:
      h->use_tma = value ? 1 : 0;
      return GXSV_OK;
    case GXSV_OPT_DIST_QUEUE:
      h->dist_queue = value ? 1 : 0;
      return GXSV_OK;
    case GXSV_OPT_TILE_IO:
      if (value < 0 || value > 2) return fail(h, GXSV_EVALUE, "tile I/O layout mode must be 0, 1 or 2");
      h->tile_io = value;
      return GXSV_OK;
    case GXSV_OPT_TILE_WARPS:
      if (value != 0 && value != 1 && value != 4 && value != 8) return fail(h, GXSV_EVALUE, "warps per tile must be 0 (auto), 1, 4 or 8");
      h->tile_warps = value;
      return GXSV_OK;
    case GXSV_OPT_DEFER_VCZ:
      h->defer_partner_cz = value != 0;
      return GXSV_OK;
    case GXSV_OPT_TILE_REGS:
      if (value != 3 && value != 4) return fail(h, GXSV_EVALUE, "tile register axes must be 3 or 4");
      { int rc = flush_batch(h); if (rc) return rc; }
      h->tile_regs = value;
      return GXSV_OK;
    default: return fail(h, GXSV_EVALUE, "unknown option %d", key);
  }
}

int gxsv_get_option(gxsv_t* h, int key, int* value) {
  switch (key) {
    case GXSV_OPT_FUSION: *value = h->fusion; return GXSV_OK;
    case GXSV_OPT_STRICT: *value = h->strict; return GXSV_OK;
    case GXSV_OPT_PROFILE: *value = h->profile; return GXSV_OK;
    case GXSV_OPT_DIST: *value = h->dist; return GXSV_OK;
    case GXSV_OPT_BATCH: *value = h->batch_max; return GXSV_OK;
    case GXSV_OPT_BATCH_AXES: *value = h->batch_axes; return GXSV_OK;
    case GXSV_OPT_CHUNK_LOG2: *value = h->chunk_log2; return GXSV_OK;
    case GXSV_OPT_TMA: *value = h->use_tma; return GXSV_OK;
    case GXSV_OPT_TILE_REGS: *value = h->tile_regs; return GXSV_OK;
    case GXSV_OPT_DIST_QUEUE: *value = h->dist_queue; return GXSV_OK;
    case GXSV_OPT_TILE_IO: *value = h->tile_io; return GXSV_OK;
    case GXSV_OPT_TILE_WARPS: *value = h->tile_warps; return GXSV_OK;
    case GXSV_OPT_DEFER_VCZ: *value = h->defer_partner_cz ? 1 : 0; return GXSV_OK;
    default: return fail(h, GXSV_EVALUE, "unknown option %d", key);
  }
}

int gxsv_nqubit(gxsv_t* h) { return (int)h->q.size(); }
int gxsv_max_qubits(gxsv_t* h) { return h->cap_bits; }

int gxsv_ensure_capacity(gxsv_t* h, int required_qubits) {
  if (required_qubits <= h->cap_bits) return GXSV_OK;
  if (required_qubits > 40) return fail(h, GXSV_EVALUE, "capacity of %d qubits exceeds the 40-qubit index space", required_qubits);
  if (!h->own_buf) return fail(h, GXSV_ENOMEM, "external buffer of 2^%d amplitudes cannot grow to 2^%d", h->cap_bits, required_qubits);
  CU(cudaSetDevice(h->device));
  { int rc = flush_batch(h); if (rc) return rc; }
  double2* nb = nullptr;
  CU(cudaMalloc(&nb, sizeof(double2) << required_qubits));
  CU(cudaMemcpyAsync(nb, h->psi, sizeof(double2) << h->phys_bits, cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaFree(h->psi));
  h->psi = nb;
  h->cap_bits = required_qubits;
  return GXSV_OK;
}

int gxsv_add_plus(gxsv_t* h) {
  CU(cudaSetDevice(h->device));
  const double r = 1.0 / std::sqrt(2.0);
  return add_one(h, cplx(r, 0.0), cplx(r, 0.0));
}

int gxsv_add_qubit(gxsv_t* h, const double amp[4]) {
  CU(cudaSetDevice(h->device));
  const cplx s0(amp[0], amp[1]), s1(amp[2], amp[3]);
  if (std::abs(s0) == 0.0 && std::abs(s1) == 0.0) return fail(h, GXSV_EVALUE, "zero single-qubit state");
  return add_one(h, s0, s1);
}

int gxsv_tensor(gxsv_t* h, int k, const double* vec) {
  CU(cudaSetDevice(h->device));
  if (k < 0 || k > 30) return fail(h, GXSV_EVALUE, "tensor: k = %d out of range", k);
  if (k == 0) { h->hscale *= cplx(vec[0], vec[1]); return GXSV_OK; }
  int rc = flush_batch(h);
  if (rc) return rc;
  rc = flush_pending(h);
  if (rc) return rc;
  const int n = nlive_bits(h);
  const uint64_t nnew = 1ull << k;
  if (n == 0) {
    // empty register: upload in canonical order (new qubit j <-> physical bit k-1-j), psi[0] (the scalar) folded in
    rc = gxsv_ensure_capacity(h, k);
    if (rc) return rc;
    // the 0-qubit state is the scalar g*hscale*psi[0]; keep it (phase included) in hscale
    double2 s0;
    Ctrl c;
    CU(cudaMemcpyAsync(&s0, h->psi, sizeof s0, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(&c, h->ctrl, sizeof c, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->hscale *= cplx(s0.x, s0.y) * c.g;
    c.g = 1.0;
    c.ticket = 0;
    CU(cudaMemcpyAsync(h->ctrl, &c, sizeof c, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->psi, vec, sizeof(double2) * nnew, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->stats.launches[GXSV_K_TENSOR] += 1;
    h->stats.algo_bytes[GXSV_K_TENSOR] += 16.0 * (double)nnew;
    h->phys_bits = k;
    h->live_mask = nnew - 1ull;
    for (int j = 0; j < k; ++j) h->q.push_back({h->next_id++, k - 1 - j, 0.0, 0.0});
    // the previous scalar psi[0]*g*hscale keeps living in hscale/g (psi[0] was 1 by construction after reset/M)
    return GXSV_OK;
  }
  NewBits nb;
  nb.k = k;
  uint64_t mask = h->live_mask;
  int extent = h->phys_bits;
  for (int j = 0; j < k; ++j) {
    int bit = -1;
    for (int b = 0; b < extent; ++b) if (!((mask >> b) & 1ull)) { bit = b; break; }
    if (bit < 0) bit = extent;
    if (bit >= extent) extent = bit + 1;
    mask |= 1ull << bit;
    nb.pos[j] = (unsigned char)bit;
  }
  if (extent > h->cap_bits) { rc = gxsv_ensure_capacity(h, extent); if (rc) return rc; }
  double2* dvec = nullptr;
  CU(cudaMalloc(&dvec, sizeof(double2) * nnew));
  CU(cudaMemcpyAsync(dvec, vec, sizeof(double2) * nnew, cudaMemcpyHostToDevice, h->stream));
  const uint64_t nlive = 1ull << n;
  Holes hs = make_holes(h);
  {
    Launch L(h, GXSV_K_TENSOR, 16.0 * (double)nlive * (double)(nnew + 1));
    k_tensor<<<grid_for(h, nlive << k, 1), TPB, 0, h->stream>>>(h->psi, hs, nb, dvec, nlive, n);
    k_scale_all<U_DEF><<<grid_for(h, nlive, U_DEF), TPB, 0, h->stream>>>(h->psi, hs, make_double2(vec[0], vec[1]), nlive);
  }
  rc = check_launch(h, "k_tensor");
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaFree(dvec));
  if (rc) return rc;
  h->live_mask = mask;
  h->phys_bits = extent;
  for (int j = 0; j < k; ++j) h->q.push_back({h->next_id++, (int)nb.pos[j], 0.0, 0.0});
  return GXSV_OK;
}

int gxsv_set_state(gxsv_t* h, int n, const double* vec_host) {
  int rc = gxsv_reset(h);
  if (rc) return rc;
  return gxsv_tensor(h, n, vec_host);
}

int gxsv_entangle(gxsv_t* h, int q1, int q2) {
  int rc = check_q(h, q1);
  if (rc) return rc;
  rc = check_q(h, q2);
  if (rc) return rc;
  if (q1 == q2) {
    // the reference's kernel flips every amplitude whose bits of q1 AND q2 are set (statevec.py:930-937): for equal
    // indices that is Z on the qubit
    const double z[8] = {1.0, 0.0, 0.0, 0.0, 0.0, 0.0, -1.0, 0.0};
    return gxsv_evolve_single(h, q1, z);
  }
  h->pend.push_back({h->q[q1].id, h->q[q2].id});
  if (h->fusion == 0) { CU(cudaSetDevice(h->device)); return flush_pending(h); }
  return GXSV_OK;
}

int gxsv_flush(gxsv_t* h) {
  CU(cudaSetDevice(h->device));
  return flush_pending(h);
}

int gxsv_evolve_single(gxsv_t* h, int q, const double m[8]) {
  CU(cudaSetDevice(h->device));
  int rc = check_q(h, q);
  if (rc) return rc;
  const cplx m00(m[0], m[1]), m01(m[2], m[3]), m10(m[4], m[5]), m11(m[6], m[7]);
  const bool diagonal = (m01 == cplx(0.0) && m10 == cplx(0.0));
  if (h->fusion >= 2) {
    Qubit& qb = h->q[q];
    if (qb.phys < 0 && (diagonal || !has_edge(h, qb.id))) {
      // still virtual: a diagonal op commutes with its queued CZs, any op does if it has none
      const cplx a = qb.s0, b = qb.s1;
      qb.s0 = m00 * a + m01 * b;
      qb.s1 = m10 * a + m11 * b;
      return GXSV_OK;
    }
    const bool is_x = (m00 == cplx(0.0) && m11 == cplx(0.0) && m01 == m10 && std::abs(m01) > 0.0);
    const bool is_z = diagonal && m11 == -m00 && std::abs(m00) > 0.0;
    if (qb.phys >= 0 && h->lazy_frame && (is_x || is_z)) {
      // byproduct on a materialised qubit: goes to the Pauli frame, no pass (c * X after Z^z X^x = (-1)^z c Z^z X^(x+1))
      if (is_x) {
        rc = settle(h, q);                   // CZs queued on q do not commute with X: they go first
        if (rc) return rc;
        const uint64_t bit = 1ull << h->q[q].phys;
        if (h->fz & bit) h->hscale = -h->hscale;
        h->fx ^= bit;
        h->hscale *= m01;
      } else {
        h->fz ^= 1ull << qb.phys;
        h->hscale *= m00;
      }
      return GXSV_OK;
    }
    if (qb.phys < 0 || !diagonal) rc = settle(h, q);
  } else {
    rc = flush_pending(h);
  }
  if (rc) return rc;
  const int p = h->q[q].phys;
  const int n = nlive_bits(h);
  const uint64_t npair = 1ull << (n - 1);
  const double S = state_bytes(h);
  Holes hs = make_holes(h, p);
  const int grid = grid_for(h, npair, U_DEF);
  if (diagonal && std::abs(m00) > 0.0) {
    // diagonal: fold m00 into the host scalar, scale only the bit = 1 half
    const cplx f = m11 / m00;
    h->hscale *= m00;
    if (f == cplx(1.0, 0.0)) return GXSV_OK;
    Launch L(h, GXSV_K_APPLY1Q, (p == 0) ? 2.0 * S : S);
    k_scale_half<U_DEF><<<grid, TPB, 0, h->stream>>>(h->psi, hs, p, d2(f), npair);
  } else {
    Launch L(h, GXSV_K_APPLY1Q, 2.0 * S);
    k_apply1q<U_DEF><<<grid, TPB, 0, h->stream>>>(h->psi, hs, p, d2(m00), d2(m01), d2(m10), d2(m11), npair);
  }
  return check_launch(h, "k_apply1q");
}

int gxsv_swap(gxsv_t* h, int q1, int q2) {
  int rc = check_q(h, q1);
  if (rc) return rc;
  rc = check_q(h, q2);
  if (rc) return rc;
  // a relabel: the amplitudes do not move (pending CZs are stored by physical bit, so they stay valid)
  std::swap(h->q[q1], h->q[q2]);
  // but queued CZs refer to physical bits of the *qubits*, which just traded places logically, not physically: nothing to fix
  return GXSV_OK;
}

int gxsv_evolve(gxsv_t* h, int k, const int* qubits, const double* op) {
  CU(cudaSetDevice(h->device));
  if (k < 1 || k > 4) return fail(h, GXSV_EUNSUPPORTED, "evolve: operators on %d qubits are not supported (1..4)", k);
  for (int i = 0; i < k; ++i) {
    int rc = check_q(h, qubits[i]);
    if (rc) return rc;
    for (int j = 0; j < i; ++j)
      if (qubits[i] == qubits[j]) return fail(h, GXSV_EVALUE, "evolve: repeated qubit %d", qubits[i]);
  }
  if (k == 1) return gxsv_evolve_single(h, qubits[0], op);
  for (int i = 0; i < k; ++i) {
    int rc = settle(h, qubits[i]);
    if (rc) return rc;
  }
  const int n = nlive_bits(h);
  const int D = 1 << k;
  // operator index bit (k-1-i) <-> qubits[i]  (op reshaped (2,)*2k with the first k legs in `qubits` order)
  GroupOffsets go;
  for (int b = 0; b < D; ++b) {
    uint64_t off = 0;
    for (int i = 0; i < k; ++i)
      if ((b >> (k - 1 - i)) & 1) off |= 1ull << h->q[qubits[i]].phys;
    go.off[b] = off;
  }
  Holes hs;
  hs.n = 0;
  for (int b = 0; b < h->phys_bits; ++b) {
    bool skip = !((h->live_mask >> b) & 1ull);
    for (int i = 0; i < k; ++i) skip = skip || (h->q[qubits[i]].phys == b);
    if (skip) hs.pos[hs.n++] = (unsigned char)b;
  }
  // the scratch is rewritten by the next evolve: stream order keeps this launch's copy intact, and the
  // (pageable) host source is consumed before cudaMemcpyAsync returns
  double2* dop = h->op_dev;
  CU(cudaMemcpyAsync(dop, op, sizeof(double2) * D * D, cudaMemcpyHostToDevice, h->stream));
  const uint64_t ngroups = 1ull << (n - k);
  const int grid = grid_for(h, ngroups, 1);
  {
    Launch L(h, GXSV_K_APPLY1Q, 2.0 * state_bytes(h));
    if (k == 2) k_evolve<2><<<grid, TPB, 0, h->stream>>>(h->psi, hs, go, dop, ngroups);
    else if (k == 3) k_evolve<3><<<grid, TPB, 0, h->stream>>>(h->psi, hs, go, dop, ngroups);
    else k_evolve<4><<<grid, TPB, 0, h->stream>>>(h->psi, hs, go, dop, ngroups);
  }
  return check_launch(h, "k_evolve");
}

int gxsv_expectation_single(gxsv_t* h, int q, const double m[8], double out[2]) {
  CU(cudaSetDevice(h->device));
  int rc = check_q(h, q);
  if (rc) return rc;
  if (h->fx | h->fz) {
    rc = flush_batch(h);
    if (!rc) rc = settle_frame(h);
    if (rc) return rc;
  }
  if (h->fusion >= 2 && h->q[q].phys < 0) {
    rc = materialize(h, q);   // the measured qubit itself must exist; its CZ partners need not
    if (rc) return rc;
  }
  if (h->fusion >= 2) {
    // Lazy mode, probability of a projector |v><v| (what `measure` asks for): nothing else has to be materialised.
    // Queued CZs between q and real qubits are a sign on the x_q = 1 amplitude; a queued CZ with a virtual partner in
    // state (s0, s1) contributes the sign (-1)^a with probability |s_a|^2, so
    //   p = w_even * sum |t0 + t1|^2 + w_odd * sum |t0 - t1|^2,   w = parity distribution of the virtual partners.
    const cplx m00(m[0], m[1]), m01(m[2], m[3]), m10(m[4], m[5]), m11(m[6], m[7]);
    const double r0 = std::norm(m00) + std::norm(m01), r1 = std::norm(m10) + std::norm(m11);
    const bool herm = std::abs(m01 - std::conj(m10)) <= 1e-14 && std::abs(m00.imag()) <= 1e-14 && std::abs(m11.imag()) <= 1e-14;
    const bool rank1 = std::abs(m00 * m11 - m01 * m10) <= 1e-13 * (r0 + r1);
    const bool unit_trace = std::abs(m00.real() + m11.real() - 1.0) <= 1e-12;
    if (herm && rank1 && unit_trace) {
      dedupe_pending(h);
      const int qid = h->q[q].id;
      uint64_t mq = 0;
      double w_even = 1.0, w_odd = 0.0;
      bool ok = true;
      for (auto& e : h->pend) {
        if (e.first != qid && e.second != qid) continue;
        Qubit* o = by_id(h, e.first == qid ? e.second : e.first);
        if (o->phys >= 0) { mq |= 1ull << o->phys; continue; }
        const double p0v = std::norm(o->s0), p1v = std::norm(o->s1);
        if (std::abs(p0v + p1v - 1.0) > 1e-12) { ok = false; break; }
        const double we = w_even * p0v + w_odd * p1v, wo = w_even * p1v + w_odd * p0v;
        w_even = we;
        w_odd = wo;
      }
      if (ok) {
        const int big = (r0 >= r1) ? 0 : 1;
        const double nb = std::sqrt(big ? r1 : r0);
        const cplx c0 = (big ? m10 : m00) / nb, c1 = (big ? m11 : m01) / nb;   // the bra <v| up to a phase
        if (!h->dist && h->analytic_norms && std::abs(w_even - w_odd) <= 1e-15 &&
            std::abs(std::norm(c0) - std::norm(c1)) <= 1e-14) {
          // one of the queued CZ partners is a not-yet-materialised qubit with |s0|^2 = |s1|^2 (every |+> of an (N, E, M)
          // stream) and the measurement is in the XY plane: p = 1/2 (|t0|^2 + |t1|^2 summed) = |c0|^2 N0 + |c1|^2 N1 = 1/2
          // for ANY normalised state — no pass over the register, and no flush of the queued projections
          out[0] = 0.5;
          out[1] = 0.0;
          h->stats.launches[GXSV_K_EXPECT] += 0;
          return GXSV_OK;
        }
        const int p = h->q[q].phys;
        const uint64_t npair = 1ull << (nlive_bits(h) - 1);
        Holes hs = make_holes(h, p);
        unsigned long long seq;
        Res* slot = next_slot(h, &seq);
        {
          Launch L(h, GXSV_K_EXPECT, state_bytes(h));
          k_pair_reduce<U_DEF, 2><<<grid_for(h, npair, U_DEF), TPB, 0, h->stream>>>(
              h->psi, hs, p, d2(c0), d2(c1), d2(0.0), d2(0.0), npair, h->partials, h->ctrl, slot, seq,
              std::norm(h->hscale), mq);
        }
        rc = check_launch(h, "k_pair_reduce");
        if (rc) return rc;
        double v[RES_VALUES];
        rc = wait_slot(h, seq, v);
        if (rc) return rc;
        out[0] = w_even * v[0] + w_odd * v[1];
        out[1] = 0.0;
        return GXSV_OK;
      }
    }
  }
  rc = settle(h, q);  // CZs not touching q and virtual qubits not entangled with q do not change <op_q>
  if (rc) return rc;
  const int p = h->q[q].phys;
  const int n = nlive_bits(h);
  const uint64_t npair = 1ull << (n - 1);
  Holes hs = make_holes(h, p);
  unsigned long long seq;
  Res* slot = next_slot(h, &seq);
  {
    Launch L(h, GXSV_K_EXPECT, state_bytes(h));
    k_pair_reduce<U_DEF, 0><<<grid_for(h, npair, U_DEF), TPB, 0, h->stream>>>(
        h->psi, hs, p, make_double2(m[0], m[1]), make_double2(m[2], m[3]), make_double2(m[4], m[5]),
        make_double2(m[6], m[7]), npair, h->partials, h->ctrl, slot, seq, std::norm(h->hscale), 0ull);
  }
  rc = check_launch(h, "k_pair_reduce");
  if (rc) return rc;
  double v[RES_VALUES];
  rc = wait_slot(h, seq, v);
  if (rc) return rc;
  out[0] = v[0];
  out[1] = v[1];
  return GXSV_OK;
}

int gxsv_project_qubit(gxsv_t* h, int q, const double m[8], double atol, double norms[2]) {
  CU(cudaSetDevice(h->device));
  int rc = check_q(h, q);
  if (rc) return rc;
  if (h->fx | h->fz) {
    // a projection may be queued without any launch: pending byproducts were issued earlier and must act first
    rc = flush_batch(h);
    if (!rc) rc = settle_frame(h);
    if (rc) return rc;
  }
  const cplx m00(m[0], m[1]), m01(m[2], m[3]), m10(m[4], m[5]), m11(m[6], m[7]);
  const double r0 = std::norm(m00) + std::norm(m01), r1 = std::norm(m10) + std::norm(m11);
  if (r0 + r1 == 0.0) return fail(h, GXSV_EZERONORM, "Attempted to remove qubit %d from 0-norm statevector.", q);
  const double det = std::abs(m00 * m11 - m01 * m10);
  const bool rank1 = det <= 1e-13 * (r0 + r1);
  double v[RES_VALUES];
  unsigned long long seq;
  if (rank1) {
    // op = u w^T: row r = u_r * w.  Contract with the larger row (best conditioned); the two half-norms the
    // reference computes (statevec.py:1185) differ from it only by the known factor |u_r|^2/|u_big|^2.
    const int big = (r0 >= r1) ? 0 : 1;
    const double rbig = big ? r1 : r0;
    // materialising queued qubits moves their amplitude s0 into the host scalar: settle first, read hscale after
    if (h->fusion < 2) rc = flush_pending(h);
    else if (h->q[q].phys < 0) rc = materialize(h, (int)q);   // only q itself: its queued CZs are handled below
    if (rc) return rc;
    cplx c0 = (big ? m10 : m00) * h->hscale, c1 = (big ? m11 : m01) * h->hscale;
    if (h->fusion < 2) {
      rc = launch_contract(h, q, c0, c1, 0ull, &seq);
      if (rc) return rc;
    } else {
      // lazy mode: the CZs queued on q are applied inside the projection pass, and one virtual
      // partner of q is materialised by the same pass (k_fused)
      dedupe_pending(h);
      const int qid = h->q[q].id;
      std::vector<int> virt;  // logical indices of virtual partners
      for (auto& e : h->pend) {
        if (e.first != qid && e.second != qid) continue;
        const int other = (e.first == qid) ? e.second : e.first;
        for (size_t i = 0; i < h->q.size(); ++i)
          if (h->q[i].id == other && h->q[i].phys < 0) virt.push_back((int)i);
      }
      int vi = -1;
      if (!virt.empty()) {
        vi = virt.back();
        virt.pop_back();
        for (int i : virt) { rc = materialize(h, i); if (rc) return rc; }
        c0 = (big ? m10 : m00) * h->hscale;
        c1 = (big ? m11 : m01) * h->hscale;
      }
      const int vid = (vi >= 0) ? h->q[vi].id : -1;
      uint64_t mq = 0, mv = 0;
      std::vector<std::pair<int, int>> keep, vedges;   // vedges: CZs of v with materialised qubits (what mv encodes)
      for (auto& e : h->pend) {
        const bool tq = (e.first == qid || e.second == qid);
        const bool tv = (vid >= 0) && (e.first == vid || e.second == vid);
        if (tq && tv) continue;  // the (q, v) edge itself: consumed by the butterfly
        if (tq) {
          mq |= 1ull << by_id(h, e.first == qid ? e.second : e.first)->phys;
        } else if (tv) {
          Qubit* o = by_id(h, e.first == vid ? e.second : e.first);
          if (o->phys >= 0) { mv |= 1ull << o->phys; vedges.push_back(e); }
          else keep.push_back(e);  // virtual-virtual edges stay queued
        } else {
          keep.push_back(e);
        }
      }
      h->pend.swap(keep);
      if (vi < 0) {
        rc = launch_contract(h, q, c0, c1, mq, &seq);
        if (rc) return rc;
      } else {
        const int p = h->q[q].phys;
        const cplx s0 = h->q[vi].s0, s1 = h->q[vi].s1;
        // |+> / |-> partners (s1 = +-s0): take the scalar f = s0 * c_pivot out of the stage (Stage kinds 0 / 1 in
        // kernels.cuh).  Its phase (sharded: f itself, so that the all-reduced norms stay those of the true state)
        // stays in the host scalar, |f|^2 goes into the published norms.  A sharded state may carry a rank-dependent
        // sign on c1, and the host scalar must be rank uniform: there only c0 is ever taken out.
        const bool same = (s1 == s0), opp = (s1 == -s0);
        const int pivot = (std::abs(c0) >= std::abs(c1)) ? 0 : 1;
        const bool fast = (same || opp) && std::abs(s0) > 0.0 && !(h->dist && pivot == 1);
        const cplx a_ratio = fast ? (pivot ? c0 / c1 : c1 / c0) : cplx(0.0);
        // |a| = 1 (XY-plane measurement, |+> / |-> partner): the outcome probability is 1/2 whatever the state, i.e.
        // |row . psi|^2 = 2 |s0|^2 |row_pivot|^2 |psi|^2 — known without looking at the register
        const bool analytic = fast && h->analytic_norms && std::abs(std::norm(a_ratio) - 1.0) <= 1e-12;
        // Queue the butterfly (up to batch_max stages on up to batch_axes distinct bits run as ONE pass) in non-strict
        // mode always, and in STRICT mode whenever the outcome probability is analytic: such a projection cannot be the
        // zero-norm branch the reference raises for (statevec.py:1192-1193), and the half the reference keeps is known,
        // so nothing about it has to be waited for — error timing and results stay exactly the reference's.
        if (h->batch_max > 0 && (!h->strict || (analytic && (!h->dist || h->dist_queue)))) {
          int ax = -1;
          for (int a = 0; a < h->batch.naxes; ++a) if (h->batch.axbit[a] == p) ax = a;
          // more than three axes need a CTA tile of 2^11 amplitudes (k_fused_tile)
          const int max_axes = (nlive_bits(h) >= 8 + h->tile_regs + 1) ? h->batch_axes : std::min(h->batch_axes, 3);
          bool full = h->batch.nstages >= std::min(h->batch_max, (int)MAX_STAGES) || (ax < 0 && h->batch.naxes >= max_axes);
          if (!full && h->batch.naxes + (ax < 0 ? 1 : 0) > 3) {
            // would the batch with this stage still fit MAX_PHASES register phases?
            Batch& b = h->batch;
            b.st[b.nstages].axis = (ax < 0) ? b.naxes : ax;
            b.nstages += 1;
            full = plan_phases(b, nullptr, h->tile_regs) > h->max_phases;
            b.nstages -= 1;
          }
          if (full) {
            rc = flush_batch(h);
            if (rc) return rc;
            ax = -1;
          }
          if (ax < 0) { ax = h->batch.naxes++; h->batch.axbit[ax] = (unsigned char)p; }
          // A queued stage applies no partner signs: the CZs of v with materialised qubits stay queued (both ends are
          // materialised once the stage has run) until one of their ends is measured — there they are part of that
          // stage's mq, which costs the same two sign XORs per pair whatever its weight — or until another kernel needs
          // the register (flush_pending, one diagonal pass).  The common stage then flips one half of its inputs only.
          if (h->defer_partner_cz) {
            h->pend.insert(h->pend.end(), vedges.begin(), vedges.end());
            mv = 0;
          }
          Stage& st = h->batch.st[h->batch.nstages++];
          st.m00 = d2(s0 * c0); st.m01 = d2(s0 * c1); st.m10 = d2(s1 * c0); st.m11 = d2(-s1 * c1);
          st.mq = mq; st.mv = mv; st.axis = ax; st.qtab = st.vtab = 0; st.pad = 0;
          cplx f = 1.0;
          st.kind = 2;
          st.neg = 0;
          st.a = d2(0.0);
          if (fast) {
            f = s0 * (pivot ? c1 : c0);
            st.a = d2(a_ratio);
            st.kind = (unsigned char)pivot;
            st.neg = opp ? 1 : 0;
          }
          const int k = h->batch.nstages - 1;
          h->batch.nscale[k] = ((k > 0 && !h->dist) ? h->batch.nscale[k - 1] : 1.0) * std::norm(f);
          // analytic probability: no need to accumulate the norm of this stage
          st.acc = 1;
          double an = 0.0;
          if (analytic) {
            const cplx mp = big ? (pivot ? m11 : m10) : (pivot ? m01 : m00);
            an = 2.0 * std::norm(s0) * std::norm(mp);
            st.acc = 0;
          }
          h->q[vi].phys = p;
          h->q.erase(h->q.begin() + q);
          h->hscale = fast ? (h->dist ? f : f / std::abs(f)) : cplx(1.0);
          h->open_notes[h->batch.nstages - 1] = phase_note(h, big, r0, r1, m00, m01, m10, m11);
          h->open_notes[h->batch.nstages - 1].an = an;
          if (!h->deferred.empty() && h->seq - h->deferred.front().first > (unsigned long long)(RING - 64)) {
            rc = check_deferred(h);   // never let the ring wrap over a result that has not been checked
            if (rc) return rc;
          }
          if (norms) { norms[0] = norms[1] = NAN; }
          return GXSV_OK;
        }
        const int n = nlive_bits(h);
        const uint64_t npair = 1ull << (n - 1);
        Holes hs = make_holes(h, p);
        Res* slot = next_slot(h, &seq);
        {
          Launch L(h, GXSV_K_FUSED, 2.0 * state_bytes(h));
          k_fused<U_DEF><<<grid_for(h, npair, U_DEF), TPB, 0, h->stream>>>(
              h->psi, hs, p, d2(s0 * c0), d2(s0 * c1), d2(s1 * c0), d2(-s1 * c1), mq, mv, npair, h->partials, h->ctrl,
              slot, seq, (h->dist ? (h->strict ? FIN_CONTRACT_DIST : FIN_ONESHOT) : FIN_CONTRACT), 1.0);
        }
        rc = check_launch(h, "k_fused");
        if (rc) return rc;
        h->q[vi].phys = p;  // the new qubit takes over the measured qubit's physical bit
        h->q.erase(h->q.begin() + q);
      }
    }
    h->hscale = 1.0;
    if (!h->strict && !h->dist) {
      Deferred d;
      d.first = seq;
      d.second = 1;
      d.note[0] = phase_note(h, big, r0, r1, m00, m01, m10, m11);
      h->deferred.push_back(d);
      if (!h->deferred.empty() && h->seq - h->deferred.front().first > (unsigned long long)(RING - 64)) {
        rc = check_deferred(h);
        if (rc) return rc;
      }
      if (norms) { norms[0] = norms[1] = NAN; }
      return GXSV_OK;
    }
    if (h->dist) {
      // sharded: hand the LOCAL half-norms back; the caller all-reduces, checks and calls gxsv_set_norm.
      // Non-strict: nothing is waited for — the caller collects the norms of a whole window (gxsv_take_norms).
      if (!h->strict) {
        Deferred d;
        d.first = seq;
        d.second = 1;
        h->deferred.push_back(d);
        if (norms) { norms[0] = norms[1] = NAN; }
        return GXSV_OK;
      }
      rc = wait_slot(h, seq, v);
      if (rc) return rc;
      if (norms) { norms[0] = v[0] * (r0 / rbig); norms[1] = v[0] * (r1 / rbig); }
      return GXSV_OK;
    }
    rc = wait_slot(h, seq, v);
    if (rc) return rc;
    const double nbig = v[0];
    const double n0 = nbig * (r0 / rbig), n1 = nbig * (r1 / rbig);
    if (norms) { norms[0] = n0; norms[1] = n1; }
    int pick = 0;
    if (n0 <= atol) {
      pick = 1;
      if (n1 <= atol) return fail(h, GXSV_EZERONORM, "Attempted to remove qubit %d from 0-norm statevector.", q);
    }
    if (pick != big) {
      // the reference keeps half `pick`: same vector up to the unit phase u_pick/u_big
      const cplx ub = big ? (std::abs(m10) >= std::abs(m11) ? m10 : m11) : (std::abs(m00) >= std::abs(m01) ? m00 : m01);
      const cplx up = big ? (std::abs(m10) >= std::abs(m11) ? m00 : m01) : (std::abs(m00) >= std::abs(m01) ? m10 : m11);
      const cplx ratio = up / ub;
      if (std::abs(ratio) > 0.0) h->hscale *= ratio / std::abs(ratio);
    }
    return GXSV_OK;
  }
  // general 2x2 (not a projector): literal two-pass reference algorithm
  if (h->dist) return fail(h, GXSV_EUNSUPPORTED, "sharded states project rank-1 operators only");
  rc = settle(h, q);
  if (rc) return rc;
  const int p = h->q[q].phys;
  const int n = nlive_bits(h);
  const uint64_t npair = 1ull << (n - 1);
  Holes hs = make_holes(h, p);
  Res* slot = next_slot(h, &seq);
  {
    Launch L(h, GXSV_K_EXPECT, state_bytes(h));
    k_pair_reduce<U_DEF, 1><<<grid_for(h, npair, U_DEF), TPB, 0, h->stream>>>(
        h->psi, hs, p, d2(m00), d2(m01), d2(m10), d2(m11), npair, h->partials, h->ctrl, slot, seq,
        std::norm(h->hscale), 0ull);
  }
  rc = check_launch(h, "k_pair_reduce");
  if (rc) return rc;
  rc = wait_slot(h, seq, v);
  if (rc) return rc;
  if (norms) { norms[0] = v[0]; norms[1] = v[1]; }
  int pick = 0;
  if (v[0] <= atol) {
    pick = 1;
    if (v[1] <= atol) return fail(h, GXSV_EZERONORM, "Attempted to remove qubit %d from 0-norm statevector.", q);
  }
  rc = launch_contract(h, q, (pick ? m10 : m00) * h->hscale, (pick ? m11 : m01) * h->hscale, 0ull, &seq);
  if (rc) return rc;
  h->hscale = 1.0;
  if (h->strict) rc = wait_slot(h, seq, v);
  return rc;
}

int gxsv_permute(gxsv_t* h, const int* perm, int n) {
  if (n != (int)h->q.size()) return fail(h, GXSV_EVALUE, "Permutation has length %d, but %d qubits expected.", n, (int)h->q.size());
  std::vector<char> seen(n, 0);
  for (int i = 0; i < n; ++i) {
    if (perm[i] < 0 || perm[i] >= n || seen[perm[i]]) return fail(h, GXSV_EVALUE, "not a permutation");
    seen[perm[i]] = 1;
  }
  std::vector<Qubit> nq(n);
  for (int i = 0; i < n; ++i) nq[i] = h->q[perm[i]];
  h->q.swap(nq);
  return GXSV_OK;
}

static int gather_chunk(gxsv_t* h, double2* out, uint64_t first, uint64_t count) {
  Launch L(h, GXSV_K_GATHER, 32.0 * (double)count, /*frame_aware=*/true);
  k_gather<<<grid_for(h, count, 4), TPB, 0, h->stream>>>(h->psi, out, h->tabs_dev, first, count, d2(h->hscale), h->ctrl,
                                                          h->fx, h->fz);
  return GXSV_OK;
}

int gxsv_flatten_device(gxsv_t* h, void* out_device) {
  CU(cudaSetDevice(h->device));
  int rc = check_deferred(h);
  if (rc) return rc;
  rc = flush_pending(h);
  if (rc) return rc;
  rc = upload_tables(h);
  if (rc) return rc;
  const uint64_t total = 1ull << h->q.size();
  gather_chunk(h, (double2*)out_device, 0, total);
  return check_launch(h, "k_gather");
}

int gxsv_flatten(gxsv_t* h, double* out_host) {
  CU(cudaSetDevice(h->device));
  int rc = check_deferred(h);
  if (rc) return rc;
  rc = flush_pending(h);
  if (rc) return rc;
  rc = upload_tables(h);
  if (rc) return rc;
  const uint64_t total = 1ull << h->q.size();
  const uint64_t chunk = std::min<uint64_t>(total, 1ull << 22);
  rc = ensure_stage(h, chunk);
  if (rc) return rc;
  double2* out = (double2*)out_host;
  int buf = 0;
  for (uint64_t first = 0; first < total; first += chunk, buf ^= 1) {
    gather_chunk(h, h->stage[buf], first, chunk);
    rc = check_launch(h, "k_gather");
    if (rc) return rc;
    CU(cudaMemcpyAsync(out + first, h->stage[buf], sizeof(double2) * chunk, cudaMemcpyDeviceToHost, h->stream));
  }
  CU(cudaStreamSynchronize(h->stream));
  return GXSV_OK;
}

int gxsv_norm2(gxsv_t* h, double* out) {
  CU(cudaSetDevice(h->device));
  int rc = flush_pending(h);
  if (rc) return rc;
  const uint64_t nlive = 1ull << nlive_bits(h);
  Holes hs = make_holes(h);
  unsigned long long seq;
  Res* slot = next_slot(h, &seq);
  {
    Launch L(h, GXSV_K_EXPECT, state_bytes(h), /*frame_aware=*/true);   // |amplitudes| do not see X / Z
    k_norm2<U_DEF><<<grid_for(h, nlive, U_DEF), TPB, 0, h->stream>>>(h->psi, hs, nlive, h->partials, h->ctrl, slot, seq,
                                                                    std::norm(h->hscale));
  }
  rc = check_launch(h, "k_norm2");
  if (rc) return rc;
  double v[RES_VALUES];
  rc = wait_slot(h, seq, v);
  if (rc) return rc;
  *out = v[0];
  return GXSV_OK;
}

// <a|b> with both registers read in place through their gather tables (no canonical copies, so it works up to the
// full capacity of the device).  out[0..1] = (re, im) of the TRUE inner product (pending scalars applied).
static int inner_in_place(gxsv_t* a, gxsv_t* b, double out[2]) {
  gxsv_t* h = a;
  if (a->q.size() != b->q.size()) return fail(h, GXSV_EVALUE, "fidelity: qubit counts differ (%d vs %d)", (int)a->q.size(), (int)b->q.size());
  if (a->device != b->device) return fail(h, GXSV_EVALUE, "fidelity: states live on different devices");
  CU(cudaSetDevice(a->device));
  for (gxsv_t* s : {a, b}) {
    int rc = check_deferred(s);
    if (!rc) rc = flush_pending(s);
    if (!rc) rc = upload_tables(s);
    if (rc) { if (s != h) h->err = s->err; return rc; }
  }
  cudaStreamSynchronize(b->stream);
  Ctrl ca, cb;
  CU(cudaMemcpyAsync(&ca, a->ctrl, sizeof ca, cudaMemcpyDeviceToHost, a->stream));
  CU(cudaMemcpyAsync(&cb, b->ctrl, sizeof cb, cudaMemcpyDeviceToHost, a->stream));
  CU(cudaStreamSynchronize(a->stream));
  const uint64_t total = 1ull << a->q.size();
  unsigned long long seq;
  Res* slot = next_slot(a, &seq);
  {
    Launch L(a, GXSV_K_EXPECT, 32.0 * (double)total, /*frame_aware=*/true);
    k_inner_gather<<<grid_for(a, total, 4), TPB, 0, a->stream>>>(a->psi, a->tabs_dev, b->psi, b->tabs_dev, total,
                                                                   a->partials, a->ctrl, slot, seq, a->fx, a->fz, b->fx, b->fz);
  }
  int rc = check_launch(a, "k_inner_gather");
  if (rc) return rc;
  double v[RES_VALUES];
  rc = wait_slot(a, seq, v);
  if (rc) return rc;
  const cplx z = cplx(v[0], v[1]) * std::conj(a->hscale * ca.g) * (b->hscale * cb.g);
  out[0] = z.real();
  out[1] = z.imag();
  return GXSV_OK;
}

int gxsv_inner(gxsv_t* a, gxsv_t* b, double out[2]) { return inner_in_place(a, b, out); }

int gxsv_fidelity(gxsv_t* a, gxsv_t* b, double* out) {
  double z[2];
  int rc = inner_in_place(a, b, z);
  if (rc) return rc;
  *out = z[0] * z[0] + z[1] * z[1];
  return GXSV_OK;
}

// A second register holding a copy of the state (same layout, same pending work applied first).
int gxsv_clone(gxsv_t* h, gxsv_t** out) {
  CU(cudaSetDevice(h->device));
  int rc = check_deferred(h);
  if (!rc) rc = flush_pending(h);
  if (rc) return rc;
  gxsv_t* c = nullptr;
  rc = gxsv_create(std::max(h->phys_bits, (int)h->q.size()), h->device, (void*)h->stream, &c);
  if (rc) { h->err = g_err; return rc; }
  c->fusion = h->fusion;
  c->q = h->q;
  c->next_id = h->next_id;
  c->phys_bits = h->phys_bits;
  c->live_mask = h->live_mask;
  c->hscale = h->hscale;
  c->fx = h->fx;
  c->fz = h->fz;
  cudaError_t e = cudaMemcpyAsync(c->psi, h->psi, sizeof(double2) << h->phys_bits, cudaMemcpyDeviceToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(c->ctrl, h->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToDevice, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { gxsv_destroy(c); return fail(h, GXSV_ECUDA, "clone: %s", cudaGetErrorString(e)); }
  *out = c;
  return GXSV_OK;
}

// Sparse read-out: entries with |amp| > atol in canonical order are NOT sorted (the host sorts by index).
int gxsv_nonzero(gxsv_t* h, double atol, unsigned long long cap, unsigned long long* idx_host, double* amp_host,
                 unsigned long long* count) {
  CU(cudaSetDevice(h->device));
  int rc = check_deferred(h);
  if (!rc) rc = flush_pending(h);
  if (!rc) rc = upload_tables(h);
  if (rc) return rc;
  unsigned long long *dcount = nullptr, *didx = nullptr;
  double2* damp = nullptr;
  CU(cudaMalloc(&dcount, sizeof(unsigned long long)));
  cudaError_t e = cudaMalloc(&didx, sizeof(unsigned long long) * std::max<unsigned long long>(cap, 1));
  if (e == cudaSuccess) e = cudaMalloc(&damp, sizeof(double2) * std::max<unsigned long long>(cap, 1));
  if (e != cudaSuccess) { cudaGetLastError(); cudaFree(dcount); cudaFree(didx); return fail(h, GXSV_ENOMEM, "nonzero: no room for %llu entries", cap); }
  cudaMemsetAsync(dcount, 0, sizeof(unsigned long long), h->stream);
  const uint64_t total = 1ull << h->q.size();
  {
    Launch L(h, GXSV_K_GATHER, 16.0 * (double)total, /*frame_aware=*/true);
    k_nonzero<<<grid_for(h, total, 4), TPB, 0, h->stream>>>(h->psi, h->tabs_dev, total, d2(h->hscale), h->ctrl, atol, dcount,
                                                             didx, damp, cap, h->fx, h->fz);
  }
  rc = check_launch(h, "k_nonzero");
  unsigned long long n = 0;
  if (!rc) {
    cudaMemcpyAsync(&n, dcount, sizeof n, cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    *count = n;
    const unsigned long long m = std::min(n, cap);
    if (m) {
      cudaMemcpyAsync(idx_host, didx, sizeof(unsigned long long) * m, cudaMemcpyDeviceToHost, h->stream);
      cudaMemcpyAsync(amp_host, damp, sizeof(double2) * m, cudaMemcpyDeviceToHost, h->stream);
      cudaStreamSynchronize(h->stream);
    }
  }
  cudaFree(dcount);
  cudaFree(didx);
  cudaFree(damp);
  return rc;
}

// ---- density matrices: rho vectorised over (row qubit, column qubit) pairs -----------------------
int gxsv_dm_trace_expect(gxsv_t* h, int npairs, const int* rows, const int* cols, int target, const double m[8],
                         double out[4]) {
  CU(cudaSetDevice(h->device));
  if (2 * npairs != (int)h->q.size()) return fail(h, GXSV_EVALUE, "dm: %d pairs do not cover %d qubits", npairs, (int)h->q.size());
  if (target >= npairs) return fail(h, GXSV_EINDEX, "dm: target pair %d out of range", target);
  for (int i = 0; i < npairs; ++i) {
    int rc = check_q(h, rows[i]);
    if (rc) return rc;
    rc = check_q(h, cols[i]);
    if (rc) return rc;
  }
  int rc = flush_pending(h);
  if (rc) return rc;
  static thread_local GatherTables t;
  memset(&t, 0, sizeof t);
  int nbits = 0;
  uint64_t bitoff[64];
  for (int i = 0; i < npairs; ++i) {
    if (i == target) continue;
    bitoff[nbits++] = (1ull << h->q[rows[i]].phys) | (1ull << h->q[cols[i]].phys);
  }
  for (int byte = 0; byte < 5; ++byte)
    for (int v = 0; v < 256; ++v) {
      uint64_t p = 0;
      for (int b = 0; b < 8; ++b)
        if (((v >> b) & 1) && byte * 8 + b < nbits) p |= bitoff[byte * 8 + b];
      t.t[byte][v] = p;
    }
  CU(cudaMemcpyAsync(h->tabs_dev, &t, sizeof t, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  const uint64_t nx = 1ull << nbits;
  uint64_t roff = 0, coff = 0;
  double2 z = make_double2(0.0, 0.0);
  double2 m00 = z, m01 = z, m10 = z, m11 = z;
  if (target >= 0) {
    roff = 1ull << h->q[rows[target]].phys;
    coff = 1ull << h->q[cols[target]].phys;
    m00 = make_double2(m[0], m[1]); m01 = make_double2(m[2], m[3]);
    m10 = make_double2(m[4], m[5]); m11 = make_double2(m[6], m[7]);
  }
  unsigned long long seq;
  Res* slot = next_slot(h, &seq);
  {
    Launch L(h, GXSV_K_EXPECT, 16.0 * (double)nx * (target >= 0 ? 4.0 : 1.0));
    k_dm_diag<<<grid_for(h, nx, 1), TPB, 0, h->stream>>>(h->psi, h->tabs_dev, nx, target >= 0 ? 1 : 0, roff, coff, m00,
                                                         m01, m10, m11, h->partials, h->ctrl, slot, seq);
  }
  rc = check_launch(h, "k_dm_diag");
  if (rc) return rc;
  double v[RES_VALUES];
  rc = wait_slot(h, seq, v);
  if (rc) return rc;
  for (int i = 0; i < 4; ++i) out[i] = v[i];
  return GXSV_OK;
}

// ---- sharded-state plumbing ---------------------------------------------------
int gxsv_nreal(gxsv_t* h) { return nlive_bits(h); }

int gxsv_is_virtual(gxsv_t* h, int q) {
  if (q < 0 || q >= (int)h->q.size()) return -1;
  return h->q[q].phys < 0 ? 1 : 0;
}

int gxsv_settle(gxsv_t* h, int q) {
  CU(cudaSetDevice(h->device));
  int rc = check_q(h, q);
  if (rc) return rc;
  if (h->fusion < 2) return flush_pending(h);
  return settle(h, q);
}

int gxsv_set_norm(gxsv_t* h, double total_norm2) {
  CU(cudaSetDevice(h->device));
  if (!(total_norm2 > 0.0)) return fail(h, GXSV_EZERONORM, "set_norm: non-positive norm");
  // total_norm2 is the norm^2 of hscale * stored over all ranks; the modulus of the host scalar moves into g so that
  // neither drifts (batched stages leave their scalar factors in hscale, see gxsv_project_qubit)
  // (sharded modes with queued projections only: there total_norm2 is by construction the norm^2 of hscale * stored)
  const bool fold = h->dist && (!h->strict || h->dist_queue) && std::abs(h->hscale) > 0.0;
  const double habs = fold ? std::abs(h->hscale) : 1.0;
  const double g = habs / std::sqrt(total_norm2);
  if (fold) h->hscale /= habs;
  CU(cudaMemcpyAsync(&h->ctrl->g, &g, sizeof g, cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return GXSV_OK;
}

// Sharded non-blocking mode: run what is queued, wait, and hand back the local norm^2 after every projection
// issued since the last call (in order).  The caller all-reduces them, checks them and calls gxsv_set_norm.
int gxsv_take_norms(gxsv_t* h, double* out, int cap, int* count) {
  CU(cudaSetDevice(h->device));
  int rc = flush_batch(h);
  if (rc) return rc;
  int n = 0;
  for (auto& d : h->deferred) {
    double v[RES_VALUES];
    rc = wait_slot(h, d.first, v);
    if (rc) return rc;
    for (int k = 0; k < d.second; ++k) {
      if (n >= cap) return fail(h, GXSV_EVALUE, "take_norms: more than %d results pending", cap);
      // a stage whose norm was not accumulated reports MINUS its analytic ratio (norm after / norm before)
      out[n++] = (d.note[k].an > 0.0) ? -d.note[k].an : v[k];
    }
  }
  h->deferred.clear();
  *count = n;
  return GXSV_OK;
}

int gxsv_get_hscale(gxsv_t* h, double out[2]) {
  out[0] = h->hscale.real();
  out[1] = h->hscale.imag();
  return GXSV_OK;
}

int gxsv_scale_host(gxsv_t* h, double re, double im) {
  h->hscale *= cplx(re, im);
  return GXSV_OK;
}

int gxsv_scale(gxsv_t* h, double re, double im) {
  CU(cudaSetDevice(h->device));
  const uint64_t nlive = 1ull << nlive_bits(h);
  Holes hs = make_holes(h);
  {
    Launch L(h, GXSV_K_MISC, 2.0 * state_bytes(h));
    k_scale_all<U_DEF><<<grid_for(h, nlive, U_DEF), TPB, 0, h->stream>>>(h->psi, hs, make_double2(re, im), nlive);
  }
  return check_launch(h, "k_scale_all");
}

static int pack_common(gxsv_t* h, int q, int bitval, uint64_t first, uint64_t count, void* buf, bool pack, double re,
                       double im) {
  CU(cudaSetDevice(h->device));
  int rc = check_q(h, q);
  if (rc) return rc;
  if (h->q[q].phys < 0 || has_edge(h, h->q[q].id))
    return fail(h, GXSV_EVALUE, "pack/unpack: qubit %d must be settled first", q);
  const int p = h->q[q].phys;
  const uint64_t half = 1ull << (nlive_bits(h) - 1);
  if (first + count > half) return fail(h, GXSV_EVALUE, "pack/unpack: range exceeds the half shard");
  Holes hs = make_holes(h, p);
  const uint64_t pv = (uint64_t)(bitval ? 1 : 0) << p;
  {
    Launch L(h, GXSV_K_MISC, 32.0 * (double)count);
    if (pack)
      k_pack_half<U_DEF><<<grid_for(h, count, U_DEF), TPB, 0, h->stream>>>(h->psi, hs, pv, first, count, (double2*)buf);
    else
      k_unpack_half<U_DEF><<<grid_for(h, count, U_DEF), TPB, 0, h->stream>>>(h->psi, hs, pv, first, count,
                                                                          (const double2*)buf, make_double2(re, im));
  }
  return check_launch(h, pack ? "k_pack_half" : "k_unpack_half");
}

int gxsv_pack_half(gxsv_t* h, int q, int bitval, uint64_t first, uint64_t count, void* out_device) {
  return pack_common(h, q, bitval, first, count, out_device, true, 1.0, 0.0);
}

int gxsv_unpack_half(gxsv_t* h, int q, int bitval, uint64_t first, uint64_t count, const void* in_device,
                     double re, double im) {
  return pack_common(h, q, bitval, first, count, (void*)in_device, false, re, im);
}

// ---- peer-memory exchange (CUDA IPC over NVLink) --------------------------------------------------------------
int gxsv_ipc_export(gxsv_t* h, unsigned char out[64]) {
  CU(cudaSetDevice(h->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t hd;
  CU(cudaIpcGetMemHandle(&hd, h->psi));
  memcpy(out, &hd, 64);
  h->exported_psi = h->psi;
  return GXSV_OK;
}

int gxsv_ipc_open(gxsv_t* h, int slot, const unsigned char handle[64]) {
  CU(cudaSetDevice(h->device));
  if (slot < 0 || slot >= 8) return fail(h, GXSV_EVALUE, "ipc_open: slot %d out of range", slot);
  if (h->peer_psi[slot]) { cudaIpcCloseMemHandle(h->peer_psi[slot]); h->peer_psi[slot] = nullptr; }
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle, 64);
  void* p = nullptr;
  CU(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
  h->peer_psi[slot] = (double2*)p;
  return GXSV_OK;
}

int gxsv_ipc_close(gxsv_t* h) {
  cudaSetDevice(h->device);
  for (int i = 0; i < 8; ++i)
    if (h->peer_psi[i]) { cudaIpcCloseMemHandle(h->peer_psi[i]); h->peer_psi[i] = nullptr; }
  return GXSV_OK;
}

// Host replay of the batched-stage arithmetic (stage_pairs) for CPU tests: v = 2^reg_bits complex amplitudes (re, im
// interleaved), updated in place.
extern "C++" {
template <int NB, int AX>
static int debug_stage_axis(double2 (&v)[1 << NB], int kind, int acc, int hasv, int uniq, const Stage& st,
                            unsigned long long tq, unsigned long long tv, double& n0, double& n1) {
  if (kind == 0 && uniq && !acc && !hasv) stage_pairs<NB, AX, 0, false, false, true>(v, st, st.a, tq, 0ull, n0, n1);
  else if (kind == 0 && !acc && !hasv) stage_pairs<NB, AX, 0, false, false>(v, st, st.a, tq, 0ull, n0, n1);
  else if (kind == 0 && acc && !hasv) stage_pairs<NB, AX, 0, true, false>(v, st, st.a, tq, 0ull, n0, n1);
  else if (kind == 0 && acc && hasv) stage_pairs<NB, AX, 0, true, true>(v, st, st.a, tq, tv, n0, n1);
  else if (kind == 1 && acc && hasv) stage_pairs<NB, AX, 1, true, true>(v, st, st.a, tq, tv, n0, n1);
  else if (kind == 2 && acc && hasv) stage_pairs<NB, AX, 2, true, true>(v, st, st.a, tq, tv, n0, n1);
  else return -1;      // not a variant the kernels instantiate
  return 0;
}
}  // extern "C++"

int gxsv_debug_stage(int reg_bits, int axis, int kind, int acc, int hasv, int uniq, double* amps, const double a[2],
                     const double m[8], unsigned long long tq, unsigned long long tv, double norm2[1]) {
  if ((reg_bits != 3 && reg_bits != 4) || axis < 0 || axis >= reg_bits) return -1;
  Stage st;
  memset(&st, 0, sizeof st);
  st.a = make_double2(a[0], a[1]);
  st.m00 = make_double2(m[0], m[1]); st.m01 = make_double2(m[2], m[3]);
  st.m10 = make_double2(m[4], m[5]); st.m11 = make_double2(m[6], m[7]);
  double n0 = 0.0, n1 = 0.0;
  int rc = -1;
  if (reg_bits == 4) {
    double2 v[16];
    for (int b = 0; b < 16; ++b) v[b] = make_double2(amps[2 * b], amps[2 * b + 1]);
    if (axis == 0) rc = debug_stage_axis<4, 0>(v, kind, acc, hasv, uniq, st, tq, tv, n0, n1);
    else if (axis == 1) rc = debug_stage_axis<4, 1>(v, kind, acc, hasv, uniq, st, tq, tv, n0, n1);
    else if (axis == 2) rc = debug_stage_axis<4, 2>(v, kind, acc, hasv, uniq, st, tq, tv, n0, n1);
    else rc = debug_stage_axis<4, 3>(v, kind, acc, hasv, uniq, st, tq, tv, n0, n1);
    for (int b = 0; b < 16; ++b) { amps[2 * b] = v[b].x; amps[2 * b + 1] = v[b].y; }
  } else {
    double2 v[8];
    for (int b = 0; b < 8; ++b) v[b] = make_double2(amps[2 * b], amps[2 * b + 1]);
    if (axis == 0) rc = debug_stage_axis<3, 0>(v, kind, acc, hasv, uniq, st, tq, tv, n0, n1);
    else if (axis == 1) rc = debug_stage_axis<3, 1>(v, kind, acc, hasv, uniq, st, tq, tv, n0, n1);
    else rc = debug_stage_axis<3, 2>(v, kind, acc, hasv, uniq, st, tq, tv, n0, n1);
    for (int b = 0; b < 8; ++b) { amps[2 * b] = v[b].x; amps[2 * b + 1] = v[b].y; }
  }
  norm2[0] = n0 + n1;
  return rc;
}

int gxsv_plan_tile_layouts(int naxes, const unsigned char* axis_bits, int nstages, const unsigned char* stage_axis,
                           unsigned long long live_mask, int reg_bits, int tile_log2, int io_mode, int* nlayouts,
                           unsigned char* s_end, unsigned char* lax, unsigned long long* gbits, unsigned long long* gtid,
                           unsigned short* sbits, unsigned short* stid) {
  if (naxes < 1 || naxes > MAX_TILE_AXES || nstages < 1 || nstages > MAX_STAGES || (reg_bits != 3 && reg_bits != 4) ||
      naxes < reg_bits || tile_log2 - reg_bits > MAX_THREAD_BITS || tile_log2 < naxes)
    return -1;
  static thread_local Batch bt;
  static thread_local TileBatch tb;
  memset(&bt, 0, sizeof bt);
  memset(&tb, 0, sizeof tb);
  bt.naxes = naxes;
  bt.nstages = nstages;
  for (int a = 0; a < naxes; ++a) bt.axbit[a] = axis_bits[a];
  for (int k = 0; k < nstages; ++k) { if (stage_axis[k] >= naxes) return -1; bt.st[k].axis = stage_axis[k]; }
  const int nph = plan_phases(bt, &tb, reg_bits, live_mask, tile_log2, io_mode);
  if (nph > MAX_PHASES) return nph;
  *nlayouts = tb.nlayouts;
  for (int p = 0; p < MAX_LAYOUTS; ++p) {
    s_end[p] = tb.s_end[p];
    for (int b = 0; b < 16; ++b) {
      gbits[p * 16 + b] = tb.gbits[p][b];
      sbits[p * 16 + b] = (unsigned short)(tb.sbits[p][b] / sizeof(double2));     // the kernel's tables hold byte offsets
    }
    for (int k = 0; k < MAX_THREAD_BITS; ++k) {
      gtid[p * MAX_THREAD_BITS + k] = tb.gtid[p][k];
      stid[p * MAX_THREAD_BITS + k] = (unsigned short)(tb.stid[p][k] / sizeof(double2));
    }
  }
  for (int k = 0; k < nstages; ++k) lax[k] = tb.lax[k];
  return nph;
}

int gxsv_flush_batch(gxsv_t* h) {
  CU(cudaSetDevice(h->device));
  int rc = flush_batch(h);
  if (!rc) rc = settle_frame(h);     // pending byproducts too: the partner is about to read this register
  return rc;
}

int gxsv_swap_peer(gxsv_t* h, int q, int slot, int my_bit, double fre, double fim) {
  CU(cudaSetDevice(h->device));
  int rc = check_q(h, q);
  if (rc) return rc;
  if (h->batch.nstages || (h->fx | h->fz))
    return fail(h, GXSV_EVALUE, "swap_peer: queued projections must be launched (gxsv_flush_batch) before the barrier that precedes the swap");
  if (slot < 0 || slot >= 8 || !h->peer_psi[slot]) return fail(h, GXSV_EVALUE, "swap_peer: no peer register in slot %d", slot);
  if (h->exported_psi != h->psi) return fail(h, GXSV_EVALUE, "swap_peer: the register moved since it was exported");
  if (h->q[q].phys < 0 || has_edge(h, h->q[q].id)) return fail(h, GXSV_EVALUE, "swap_peer: qubit %d must be settled first", q);
  const int p = h->q[q].phys;
  const uint64_t half = 1ull << (nlive_bits(h) - 1);
  const int B = my_bit ? 1 : 0;
  // the two ranks of a pair split the index range: bit 0 takes the lower part, bit 1 the upper part
  const uint64_t first = B ? half / 2 : 0, count = B ? half - half / 2 : half / 2;
  if (count == 0) return GXSV_OK;
  Holes hs = make_holes(h, p);
  const uint64_t my_or = (uint64_t)(1 - B) << p, peer_or = (uint64_t)B << p;
  const cplx f(fre, fim), finv = cplx(1.0) / f;
  const int scale = (f != cplx(1.0, 0.0)) ? 1 : 0;
  // diagnostics (tools/microbench_swap.py): GXSV_SWAP_MODE = 1 pull only / 2 push only, GXSV_SWAP_U = loads in flight per
  // thread, GXSV_SWAP_CTAS = CTAs per SM
  static const int dbg_mode = std::getenv("GXSV_SWAP_MODE") ? std::atoi(std::getenv("GXSV_SWAP_MODE")) : 0;
  static const int dbg_u = std::getenv("GXSV_SWAP_U") ? std::atoi(std::getenv("GXSV_SWAP_U")) : U_DEF;
  static const int dbg_ctas = std::getenv("GXSV_SWAP_CTAS") ? std::atoi(std::getenv("GXSV_SWAP_CTAS")) : 8;
  static const int blocked = std::getenv("GXSV_SWAP_BLOCKED") ? std::atoi(std::getenv("GXSV_SWAP_BLOCKED")) : 1;
  {
    Launch L(h, GXSV_K_MISC, 64.0 * (double)count);
    const int u = (dbg_u == 8 || dbg_u == 2 || dbg_u == 1) ? dbg_u : U_DEF;
    uint64_t blocks = (count + (uint64_t)TPB * u - 1) / ((uint64_t)TPB * u);
    blocks = std::max<uint64_t>(1, std::min<uint64_t>(blocks, (uint64_t)h->sm_count * std::max(1, dbg_ctas)));
    // contiguous range per CTA, a multiple of the per-iteration chunk
    const uint64_t cpi = (uint64_t)TPB * u;
    const uint64_t span = blocked ? ((count + blocks - 1) / blocks + cpi - 1) / cpi * cpi : 0;
    if (u == 8)
      k_swap_peer<8><<<(int)blocks, TPB, 0, h->stream>>>(h->psi, h->peer_psi[slot], hs, my_or, peer_or, first, count, d2(f), d2(finv), scale, dbg_mode, span);
    else if (u == 2)
      k_swap_peer<2><<<(int)blocks, TPB, 0, h->stream>>>(h->psi, h->peer_psi[slot], hs, my_or, peer_or, first, count, d2(f), d2(finv), scale, dbg_mode, span);
    else if (u == 1)
      k_swap_peer<1><<<(int)blocks, TPB, 0, h->stream>>>(h->psi, h->peer_psi[slot], hs, my_or, peer_or, first, count, d2(f), d2(finv), scale, dbg_mode, span);
    else
      k_swap_peer<U_DEF><<<(int)blocks, TPB, 0, h->stream>>>(h->psi, h->peer_psi[slot], hs, my_or, peer_or, first, count, d2(f), d2(finv), scale, dbg_mode, span);
  }
  return check_launch(h, "k_swap_peer");
}

int gxsv_sync(gxsv_t* h) {
  CU(cudaSetDevice(h->device));
  int rc = check_deferred(h);   // runs queued projections and raises their zero-norm errors (non-strict mode)
  CU(cudaStreamSynchronize(h->stream));
  return rc;
}

int gxsv_stats(gxsv_t* h, gxsv_stats_t* out) {
  flush_batch(h);
  prof_collect(h);
  *out = h->stats;
  return GXSV_OK;
}

int gxsv_reset_stats(gxsv_t* h) {
  prof_collect(h);
  memset(&h->stats, 0, sizeof h->stats);
  return GXSV_OK;
}

int gxsv_layout(gxsv_t* h, int* out) {
  for (size_t i = 0; i < h->q.size(); ++i) out[i] = h->q[i].phys;
  return GXSV_OK;
}

}  // extern "C"
