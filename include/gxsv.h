/*
 * gxsv.h — C ABI of the B200-native complex128 statevector engine that sits
 * under graphix_b200.B200StatevectorBackend.
 *
 * The reference (TeamGraphix/graphix) has no FFI: its statevector path is a
 * Python class (`graphix/sim/statevec.py:52-750`, class Statevector) whose
 * methods call numba kernels (`statevec.py:841-1198`).  Every entry point
 * below replaces one of those methods; the citation on each declaration is the
 * reference method (and numba kernel) whose semantics it reproduces.  The
 * Python binding a graphix maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - All functions return an int status (GXSV_OK == 0).  No C++ exception ever
 *    crosses this boundary; `gxsv_last_error` returns the message.
 *  - Qubit indices are LOGICAL and follow the reference convention
 *    (`statevec.py:899-900,965-969`): qubit 0 is the most significant bit of
 *    the flattened state, newly added qubits take the last indices.  The
 *    physical layout of the amplitudes in HBM is private to the library.
 *  - Complex numbers are pairs of doubles (re, im); 2x2 operators are 8
 *    doubles in row-major order (m00, m01, m10, m11).
 *  - One host thread drives a handle at a time (the reference's Backend
 *    contract is strictly sequential, `graphix/simulator.py:449-478`).
 *  - Device memory, streams and events are owned by the library unless an
 *    external buffer/stream is handed in at creation.
 */
#ifndef GXSV_H_
#define GXSV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gxsv gxsv_t;

/* status codes; the Python layer maps them to the reference's exception types
 * (RuntimeError / IndexError / ValueError). */
enum {
  GXSV_OK = 0,
  GXSV_EINDEX = 1,    /* qubit index out of range  -> IndexError  (statevec.py:526-542) */
  GXSV_EVALUE = 2,    /* inconsistent arguments    -> ValueError  (statevec.py:1221-1225) */
  GXSV_EZERONORM = 3, /* projection onto 0-norm    -> RuntimeError(statevec.py:1192-1193) */
  GXSV_ECUDA = 4,     /* CUDA runtime failure      -> RuntimeError */
  GXSV_ENOMEM = 5,    /* device allocation failed  -> MemoryError */
  GXSV_EUNSUPPORTED = 6
};

/* option keys for gxsv_set_option */
enum {
  GXSV_OPT_FUSION = 1,  /* 0 = one kernel per command, 1 = fuse runs of E (library default), 2 = lazy N/E->M fusion
                         * (default of the Python backend) */
  GXSV_OPT_STRICT = 2,  /* 1 (default) = project_qubit waits for the norm and raises like the reference */
  GXSV_OPT_PROFILE = 3, /* 1 = bracket every launch with CUDA events (read back through gxsv_stats) */
  GXSV_OPT_DIST = 4,    /* 1 = this handle is one shard of a distributed state (see the sharded section) */
  GXSV_OPT_BATCH = 5,   /* K > 0 (needs FUSION = 2, STRICT = 0): queue up to K fused projections and execute them in
                         * ONE pass over the register; zero-norm errors surface at gxsv_sync / gxsv_flatten */
  GXSV_OPT_BATCH_AXES = 6, /* tuning: distinct physical bits one batched pass may act on (1..7, default 7) */
  GXSV_OPT_CHUNK_LOG2 = 7, /* tuning: log2 of the consecutive blocks / tiles a CTA of a batched pass takes at a time */
  GXSV_OPT_TMA = 8,        /* 1 = batched passes stage CTA tiles through shared memory with bulk copies (cp.async.bulk +
                            * mbarrier) whenever the layout allows; 0 (default) = register-staged kernels */
  GXSV_OPT_TILE_REGS = 9,  /* tuning: register axes per phase of the multi-axis batched pass: 3 (2^3 amplitudes per thread,
                            * 4 CTAs per SM) or 4 (default: 2^4 per thread, 2 CTAs per SM, fewer phases) */
  GXSV_OPT_DIST_QUEUE = 10,/* sharded STRICT mode: 1 = projections whose outcome probability is analytic are queued like in
                            * the single-GPU strict mode (gxsv_project_qubit then returns NaN norms and the caller collects
                            * them later with gxsv_take_norms); everything else still returns its local half-norms at once */
  GXSV_OPT_TILE_IO = 11,   /* tuning: load / store layouts of the multi-axis batched pass: 1 (default) = when the three lowest bits
                            * of a tile are all batch axes the tile is loaded and stored with the lanes on its LOWEST bits (contiguous
                            * 512-byte runs) and transposed to / from the first / last phase in shared memory; 0 = never
                            * (lanes on the dense bits of the first / last phase), 2 = always */
  GXSV_OPT_TILE_WARPS = 13,/* tuning: warps that share one tile of the multi-axis batched pass: 0 (default) = by the number
                            * of axes (4 from 6 axes on: runs of >= 256 bytes), 1 = private warp tiles, 4, 8 */
  GXSV_OPT_DEFER_VCZ = 12  /* tuning: 1 (default) = a queued projection leaves the CZs of the newly materialised partner
                            * queued (they are applied, for free, by the projection that measures one of their ends);
                            * 0 = they are applied as signs inside the stage that materialises the partner */
};

/* kernel kinds reported by gxsv_profile_read / gxsv_stats */
enum {
  GXSV_K_FILL = 0,     /* N: append one qubit (copy-fill of a free physical bit) */
  GXSV_K_CZ = 1,       /* E-run: fused diagonal CZ phase pass */
  GXSV_K_CONTRACT = 2, /* M: project + compact + norm reduction */
  GXSV_K_APPLY1Q = 3,  /* X / Z / C: 2x2 on one axis */
  GXSV_K_EXPECT = 4,   /* probability / expectation reduction */
  GXSV_K_GATHER = 5,   /* finalize / flatten permutation gather */
  GXSV_K_FUSED = 6,    /* lazy N,E..,M fused butterfly */
  GXSV_K_TENSOR = 7,   /* general add_nodes (2^k vector) */
  GXSV_K_MISC = 8,
  GXSV_K_COUNT = 9
};

typedef struct {
  uint64_t launches[GXSV_K_COUNT];   /* kernels launched per kind since create/reset_stats */
  double algo_bytes[GXSV_K_COUNT];   /* algorithmic bytes of those launches */
  double device_ms[GXSV_K_COUNT];    /* CUDA-event time per kind (only while GXSV_OPT_PROFILE = 1) */
} gxsv_stats_t;

/* ---- lifetime ------------------------------------------------------- */

/* Statevector(nqubit=0, max_qubits=N)  — statevec.py:93-216, with_capacity :807-832.
 * `stream` is a cudaStream_t (NULL = legacy default stream). */
int gxsv_create(int max_qubits, int device, void* stream, gxsv_t** out);
/* Same, on caller-owned device memory of 2^cap_qubits complex128 (used for sharded states). */
int gxsv_create_external(void* device_buffer, int cap_qubits, int device, void* stream, gxsv_t** out);
int gxsv_destroy(gxsv_t* h);
/* back to the 0-qubit scalar state [1] */
int gxsv_reset(gxsv_t* h);
const char* gxsv_last_error(gxsv_t* h);
int gxsv_set_option(gxsv_t* h, int key, int value);
int gxsv_get_option(gxsv_t* h, int key, int* value);

/* ---- queries -------------------------------------------------------- */

int gxsv_nqubit(gxsv_t* h);       /* Statevector.nqubit      statevec.py:232-236 */
int gxsv_max_qubits(gxsv_t* h);   /* Statevector.max_qubits  statevec.py:238-241 */
/* Statevector.ensure_capacity    statevec.py:243-259 */
int gxsv_ensure_capacity(gxsv_t* h, int required_qubits);

/* ---- N: add qubits -------------------------------------------------- */

/* add_nodes fast path, |+> appended as the new last qubit
 * statevec.py:269-300 -> _add_default_node_jit :867-887 */
int gxsv_add_plus(gxsv_t* h);
/* one qubit in state (s0, s1) = amp[0..3]; PlanarState.to_statevector_numpy states.py:67-79 */
int gxsv_add_qubit(gxsv_t* h, const double amp[4]);
/* psi <- psi (x) vec, vec = 2^k complex128 in reference order
 * Statevector.tensor statevec.py:508-524 -> _tensor_jit :841-864 */
int gxsv_tensor(gxsv_t* h, int k, const double* vec);

/* ---- E: CZ ---------------------------------------------------------- */

/* Statevector.entangle statevec.py:302-315 -> _entangle_jit :915-937.
 * With GXSV_OPT_FUSION >= 1 the CZ is queued and consecutive calls are fused
 * into one diagonal phase pass at the next non-E call. */
int gxsv_entangle(gxsv_t* h, int q1, int q2);
/* force queued work onto the device */
int gxsv_flush(gxsv_t* h);

/* ---- X / Z / C: single-qubit 2x2 ------------------------------------ */

/* Statevector.evolve_single statevec.py:317-332 -> _evolve_single_jit :943-986 */
int gxsv_evolve_single(gxsv_t* h, int q, const double m[8]);
/* Statevector.swap statevec.py:497-506 -> _swap_jit :890-912 (a relabel here) */
int gxsv_swap(gxsv_t* h, int q1, int q2);
/* Statevector.evolve statevec.py:361-395, k-qubit operator (2^k x 2^k row-major) */
int gxsv_evolve(gxsv_t* h, int k, const int* qubits, const double* op);

/* ---- M: probability and projection ---------------------------------- */

/* Statevector.expectation_single statevec.py:334-359 -> _expectation_single_jit :992-1039
 * out[0..1] = <psi| op_q |psi> (complex). Synchronous. */
int gxsv_expectation_single(gxsv_t* h, int q, const double m[8], double out[2]);
/* Statevector.project_qubit statevec.py:430-495 -> _project_qubit_jit :1160-1198:
 * apply `m` on qubit q, keep the non-zero half, normalize, remove the qubit.
 * GXSV_EZERONORM when both half-norms are <= atol.  norms (may be NULL)
 * receives the two half-norms the reference computes (:1185). */
int gxsv_project_qubit(gxsv_t* h, int q, const double m[8], double atol, double norms[2]);

/* ---- finalize / read-out -------------------------------------------- */

/* Statevector.permute statevec.py:544-550: new qubit i is old qubit perm[i] */
int gxsv_permute(gxsv_t* h, const int* perm, int n);
/* Statevector.flatten statevec.py:261-267: 2^nqubit complex128, qubit 0 = MSB, to HOST memory */
int gxsv_flatten(gxsv_t* h, double* out_host);
/* same, to DEVICE memory owned by the caller (2^nqubit complex128) */
int gxsv_flatten_device(gxsv_t* h, void* out_device);
/* |<a|b>|^2 on device — Statevector.fidelity statevec.py:552-568; both registers are read in place through their
 * own layouts (no canonical copies) */
int gxsv_fidelity(gxsv_t* a, gxsv_t* b, double* out);
/* <a|b> (re, im) — the inner product Statevector.expectation_value ends with (statevec.py:397-415) */
int gxsv_inner(gxsv_t* a, gxsv_t* b, double out[2]);
/* a second register holding a copy of the state — the deepcopy of Statevector.expectation_value (statevec.py:411) */
int gxsv_clone(gxsv_t* h, gxsv_t** out);
/* sparse read-out for Statevector.to_dict / to_prob_dict (statevec.py:591-678): canonical indices (qubit 0 = MSB) and
 * amplitudes of the entries with |amp| > atol, in no particular order; *count = how many there are (may exceed cap: then
 * only cap were stored) */
int gxsv_nonzero(gxsv_t* h, double atol, unsigned long long cap, unsigned long long* idx_host, double* amp_host,
                 unsigned long long* count);
/* <psi|psi> of the live state (1 for a normalized state) */
int gxsv_norm2(gxsv_t* h, double* out);
/* replace the state by 2^n host amplitudes in reference order (Statevector(data), statevec.py:93-216) */
int gxsv_set_state(gxsv_t* h, int n, const double* vec_host);

/* ---- sharded states (no counterpart in the single-process reference) ----
 * With GXSV_OPT_DIST = 1 the handle holds the local sub-cube of a state sharded over ranks by its top
 * (rank-index) qubits.  gxsv_project_qubit then returns the LOCAL half-norms and leaves the
 * normalisation to gxsv_set_norm(total) after the caller's all-reduce; gxsv_expectation_single
 * returns the local contribution.  pack/unpack move one half of the local sub-cube (bit of qubit q
 * = bitval, dense order, elements [first, first+count)) to / from a contiguous device staging
 * buffer for the half-shard exchange that swaps a rank-index qubit with local qubit q. */
int gxsv_nreal(gxsv_t* h);                 /* materialised local qubits */
int gxsv_is_virtual(gxsv_t* h, int q);     /* 1 = prepared but not materialised (lazy mode) */
int gxsv_settle(gxsv_t* h, int q);         /* materialise q and issue every queued CZ touching it */
int gxsv_set_norm(gxsv_t* h, double total_norm2);
/* non-blocking sharded projections (STRICT = 0, optionally BATCH = K): every projection pass consumes the pending
 * factor g once and leaves g = 1; gxsv_take_norms flushes, waits and returns the local norm^2 after each projection
 * issued since the previous call, so that the caller can all-reduce, check and renormalise a whole window at once. */
int gxsv_take_norms(gxsv_t* h, double* out, int cap, int* count);
int gxsv_scale_host(gxsv_t* h, double re, double im); /* multiply the pending host scalar (no pass) */
int gxsv_get_hscale(gxsv_t* h, double out[2]);        /* read the pending host scalar */
int gxsv_scale(gxsv_t* h, double re, double im);      /* multiply every local amplitude (one pass) */
int gxsv_pack_half(gxsv_t* h, int q, int bitval, uint64_t first, uint64_t count, void* out_device);
int gxsv_unpack_half(gxsv_t* h, int q, int bitval, uint64_t first, uint64_t count, const void* in_device,
                     double re, double im);

/* Peer-memory exchange (one process per GPU): every rank exports its register as a CUDA IPC handle (64 bytes), the
 * host layer all-gathers the handles and each rank opens those of the ranks it pairs with (one slot per rank-index
 * bit).  gxsv_swap_peer then swaps, IN PLACE on both GPUs through NVLink loads / stores, the half with bit(q) = 1 - my_bit
 * of this shard with the half with bit(q) = my_bit of the partner's shard (amplitudes arriving here are multiplied by
 * (fre, fim), those arriving there by its inverse); both partners call it, each handles half of the elements.  The
 * caller must place a cross-rank barrier on the stream before and after the call ("qubit-remap"). */
int gxsv_ipc_export(gxsv_t* h, unsigned char handle_out[64]);
int gxsv_ipc_open(gxsv_t* h, int slot, const unsigned char handle[64]);
int gxsv_ipc_close(gxsv_t* h);
int gxsv_swap_peer(gxsv_t* h, int q, int slot, int my_bit, double fre, double fim);
/* Host replay of ONE batched stage (the arithmetic of k_fused_batch / k_fused_tile, compiled for the CPU from the same source)
 * on the 2^reg_bits amplitudes a thread holds, for tests without a device (tests/test_stage_math.py): amps = 2^reg_bits complex
 * numbers (re, im interleaved, updated in place), axis = register axis of the stage, kind / acc / hasv / uniq = the variant
 * (kernels.cuh: stage_pairs), a = the multiplier of kinds 0 / 1, m = the 2x2 of kind 2 (m00, m01, m10, m11), tq / tv = sign
 * tables (byte k = 0x80: the k-th pair takes the sign), norm2 = accumulated norm^2 of the results (acc = 1).  Returns -1 for a
 * variant the kernels do not instantiate. */
int gxsv_debug_stage(int reg_bits, int axis, int kind, int acc, int hasv, int uniq, double* amps, const double a[2],
                     const double m[8], unsigned long long tq, unsigned long long tv, double norm2[1]);

/* Host-side planner of the multi-axis batched pass (k_fused_tile), exposed so that its tables can be checked without a
 * device: for a batch of `nstages` fused projections whose k-th stage acts on batch axis stage_axis[k] (physical bit
 * axis_bits[stage_axis[k]]), returns the number of register phases and fills, per layout p < *nlayouts (at most 6):
 * s_end[p] (stages run in layouts 0..p), gbits[16 p + b] / sbits[16 p + b] (physical offset / shared-tile slot of register
 * slot b), gtid[8 p + i] / stid[8 p + i] (the same for bit i of a thread's index in its group), and lax[k] (register slot
 * of stage k's axis).  live_mask = live physical bits, reg_bits = 3 | 4, tile_log2 = log2 amplitudes per tile, io_mode as
 * GXSV_OPT_TILE_IO.  Returns -1 on invalid arguments; a result > 4 means the batch does not fit (tables not filled). */
int gxsv_plan_tile_layouts(int naxes, const unsigned char* axis_bits, int nstages, const unsigned char* stage_axis,
                           unsigned long long live_mask, int reg_bits, int tile_log2, int io_mode, int* nlayouts,
                           unsigned char* s_end, unsigned char* lax, unsigned long long* gbits, unsigned long long* gtid,
                           unsigned short* sbits, unsigned short* stid);

/* launch the queued batched projections now (everything this shard still has to do to its register must be on the
 * stream BEFORE the barrier that precedes a peer swap: the partner reads this register right after that barrier) */
int gxsv_flush_batch(gxsv_t* h);

/* ---- density matrices (BASELINE config 5; graphix/sim/density_matrix.py) -------------------------
 * rho over n qubits is held vectorised as a 2n-qubit register: qubit i has a row qubit rows[i] and a
 * column qubit cols[i] (logical indices of this handle).  U rho U^dagger = (U on rows[i], conj(U) on
 * cols[i]) through gxsv_evolve_single / gxsv_evolve; a Kraus channel on k qubits is one 2k-axis
 * gxsv_evolve with sum_j |c_j|^2 K_j (x) conj(K_j); a rank-1 projection is gxsv_project_qubit on the row
 * qubit and on the column qubit (GXSV_OPT_DIST = 1 so that no vector norm is imposed), followed by
 * gxsv_set_norm(Tr^2).  This entry point supplies the traces:
 *   out[0..1] = Tr(op_target rho)  (density_matrix.py:263-290; skipped when target < 0)
 *   out[2..3] = Tr(rho)            (density_matrix.py:354-365)
 * in stored units (their ratio is the normalised expectation value). */
int gxsv_dm_trace_expect(gxsv_t* h, int npairs, const int* rows, const int* cols, int target, const double m[8],
                         double out[4]);

/* ---- plumbing -------------------------------------------------------- */

int gxsv_sync(gxsv_t* h);
int gxsv_stats(gxsv_t* h, gxsv_stats_t* out);
int gxsv_reset_stats(gxsv_t* h);
/* physical bit of each logical qubit (debug / tests), out has nqubit entries; -1 = not materialized */
int gxsv_layout(gxsv_t* h, int* out);
int gxsv_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GXSV_H_ */
