/*
 * zkp_b200 -- C ABI of the B200-native KZG prover backend.
 *
 * This is the drop-in boundary for the hot path of apollozkp/zkp-subnet: everything the miner and
 * validator neurons obtain from the external Rust prover through `fourier.Client`
 * (reference base/miner.py:26,73-84; base/validator.py:28,80-91).  Each entry point names the
 * reference call site it replaces.  Conventions:
 *   - every function returns 0 on success or a negative zkp_status; no exceptions, no aborts;
 *     zkp_last_error() gives a human-readable message for the calling thread's last failure;
 *   - field elements cross the boundary as 32-byte BIG-ENDIAN canonical integers (< r), exactly the
 *     bytes inside the reference's base64 `poly` / `alpha` / `eval` strings
 *     (reference base/protocol.py:35-60, tests/test_miner.py:33-55);
 *   - G1 points cross as 48-byte ZCash-compressed encodings (the bytes inside the reference's
 *     base64 `commitment` / `proof` strings) or, for SRS import/export, 96-byte uncompressed;
 *   - the caller owns every buffer; the library never returns heap pointers;
 *   - a context is thread-safe (calls are serialised internally); several contexts may coexist
 *     in one process (miner on port 1337 + validator on port 1338 in the reference).
 * There is no CPU fallback: without a CUDA device zkp_ctx_create fails with ZKP_ERR_CUDA.
 */
#ifndef ZKP_B200_H
#define ZKP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zkp_ctx zkp_ctx;

typedef enum {
    ZKP_OK = 0,
    ZKP_ERR_ARG = -1,      /* bad argument (null pointer, size not a power of two, row out of range) */
    ZKP_ERR_ENCODING = -2, /* non-canonical field element / malformed point */
    ZKP_ERR_CUDA = -3,     /* CUDA runtime failure or no device */
    ZKP_ERR_STATE = -4,    /* SRS not loaded / wrong size */
    ZKP_ERR_IO = -5        /* file error */
} zkp_status;

/* ---- lifecycle: replaces Client(port, bin, uncompressed, setup_path, precompute_path),
 *      client.start(scale, machines_scale) and client.stop()
 *      (reference base/miner.py:73-84,155,181; base/validator.py:80-91,173,200) */
int zkp_ctx_create(int device, zkp_ctx** out);
/* A further context on the same device that SHARES the parent's resident SRS, fixed-base tables and domain tables
 * (one copy in HBM) and has its own streams and workspaces: one per request-handling thread (the reference hands
 * `forward` to the axon's threads, base/miner.py:66-70).  Destroy every fork with zkp_ctx_destroy; the SRS lives as
 * long as any context that shares it.  Do not replace the SRS while forks are computing. */
int zkp_ctx_fork(zkp_ctx* parent, zkp_ctx** out);
void zkp_ctx_destroy(zkp_ctx* ctx);
const char* zkp_last_error(void);
int zkp_device_count(void);

/* ---- pinned host staging.  A polynomial crosses the boundary as n x 32 bytes; when those bytes sit in a
 *      buffer from zkp_host_alloc (page-locked) the host->device copy inside zkp_worker_* runs as one
 *      asynchronous DMA at PCIe rate instead of being staged through the driver's bounce buffer.  Any
 *      ordinary (pageable) pointer is accepted too -- the result is identical, only the copy is slower.
 *      The Python client decodes the base64 `poly` list straight into such a buffer. */
int zkp_host_alloc(size_t bytes, void** out);
int zkp_host_free(void* p);

/* ---- SRS: replaces `prover setup --generate-setup --generate-precompute` and the
 *      setup_path / precompute_path files (reference tests/conftest.py:50-65, Makefile:30-48).
 *      Layout: 2^log_machines rows of 2^log_n points, U[i][j] = [R_i(tau_y) L_j(tau_x)]_1
 *      (Lagrange basis over the natural-order domains), plus [R_i(tau_y)]_1 per row and [tau_x]_2. */
int zkp_srs_generate(zkp_ctx* ctx, const uint8_t tau_x_be[32], const uint8_t tau_y_be[32],
                     uint32_t log_n, uint32_t log_machines);
/* the monomial SRS [tau_x^j]_1 as the single row (commitment from COEFFICIENTS; "path B" of BASELINE configs[2]:
 * zkp_fft(inverse) then zkp_msm_g1 over this row equals zkp_worker_commit over the Lagrange row, byte for byte) */
int zkp_srs_generate_monomial(zkp_ctx* ctx, const uint8_t tau_x_be[32], uint32_t log_n);
/* point-range shard `shard` of 2^log_shards of the same SRS: rows hold points j in
 * [shard*n/S, (shard+1)*n/S); an MSM over the row with the matching scalar slice is one GPU's partial
 * commitment (MSM sharding by point range) */
int zkp_srs_generate_shard(zkp_ctx* ctx, const uint8_t tau_x_be[32], const uint8_t tau_y_be[32],
                           uint32_t log_n, uint32_t log_machines, uint32_t shard, uint32_t log_shards);
/* Opening of a polynomial split by point range over G GPUs, each holding the shard made by
 * zkp_srs_generate_shard and the matching slice of evaluations.  The barycentric sum for y = f(x) splits by
 * point range exactly like the MSM: (1) every rank computes its partial sum (32 bytes) with
 * zkp_shard_eval_partial, (2) the G partials are gathered and turned into y on the host by
 * zkp_shard_eval_combine (log_n = the FULL domain), (3) every rank computes the partial proof of its slice
 * with zkp_shard_open_partial; proof = zkp_g1_sum of the G partial proofs.  One 32-byte and one 48-byte
 * exchange per opening, no collective inside any kernel.  x inside the evaluation domain is refused
 * (ZKP_ERR_ARG) in this mode -- it needs a second global sum and has probability 2^-235 for a random x. */
int zkp_shard_eval_partial(zkp_ctx* ctx, uint32_t i, const uint8_t* slice_be, size_t n_local, const uint8_t x_be[32],
                           uint8_t partial_be[32]);
int zkp_shard_eval_combine(const uint8_t* partials_be, size_t count, uint32_t log_n, const uint8_t x_be[32],
                           uint8_t y_be[32]);
int zkp_shard_open_partial(zkp_ctx* ctx, uint32_t i, const uint8_t* slice_be, size_t n_local, const uint8_t x_be[32],
                           const uint8_t y_be[32], uint8_t proof_partial48[48]);
/* sum of `count` compressed G1 points: the cross-GPU combine of partial commitments / Pianist
 * aggregation com = sum_i com_i, pi = sum_i pi_i (host arithmetic, 48 bytes per GPU) */
int zkp_g1_sum(const uint8_t* points48, size_t count, uint8_t out48[48]);
/* the same for points received from other parties (the master node aggregating workers' answers): every input is
 * checked to be on the curve AND in the prime-order subgroup; ZKP_ERR_ENCODING otherwise */
int zkp_g1_sum_checked(const uint8_t* points48, size_t count, uint8_t out48[48]);
/* the same with 96-byte ZCash-uncompressed inputs (no square roots on the combining rank); zkp_g1_uncompress
 * expands a compressed point on the rank that produced it */
int zkp_g1_uncompress(const uint8_t in48[48], uint8_t out96[96]);
/* commitment || proof of the last commit+open on this context, 2 x 96 bytes uncompressed (no square root anywhere in
 * a cross-process combine: every rank contributes these, rank 0 adds them with zkp_g1_sum_uncompressed) */
int zkp_last_points_uncompressed(zkp_ctx* ctx, uint8_t out192[192]);
/* the same as two JACOBIAN points (X, Y, Z: 3 x 48 bytes of this library's Montgomery limbs each; Z = 0 = infinity): no
 * field inversion on the contributing rank at all.  zkp_g1_sum_jacobian adds `count` such records (`stride` >= 144
 * bytes apart; each checked to be on the curve) and compresses the sum.  Internal representation: only for exchange
 * between processes of the same build on one box (the multi-process combine of bench.py). */
int zkp_last_points_jacobian(zkp_ctx* ctx, uint8_t out288[288]);
int zkp_g1_sum_jacobian(const uint8_t* points, size_t count, size_t stride, uint8_t out48[48]);
int zkp_g1_sum_uncompressed(const uint8_t* points96, size_t count, uint8_t out48[48]);
/* import one row from 96-byte ZCash-uncompressed points (validated on curve), and its scale point */
int zkp_srs_set_shape(zkp_ctx* ctx, uint32_t log_n, uint32_t log_machines);
int zkp_srs_import_row(zkp_ctx* ctx, uint32_t row, const uint8_t* points96, size_t n,
                       const uint8_t scale_point48[48]);
int zkp_srs_import_g2_tau(zkp_ctx* ctx, const uint8_t tau_x_be[32]);
int zkp_srs_import_g2_tau_y(zkp_ctx* ctx, const uint8_t tau_y_be[32]); /* [tau_y]_2, master verification only */
int zkp_srs_export_row(zkp_ctx* ctx, uint32_t row, uint8_t* points96, size_t n);
/* the same with 48-byte ZCash-COMPRESSED points (the reference's `.compressed` files, tests/conftest.py:28-29;
 * `--uncompressed false`): square roots and curve checks run on the device */
int zkp_srs_import_row_compressed(zkp_ctx* ctx, uint32_t row, const uint8_t* points48, size_t n, const uint8_t scale_point48[48]);
int zkp_srs_export_row_compressed(zkp_ctx* ctx, uint32_t row, uint8_t* points48, size_t n);
/* [tau_x]_2 (which = 0) and [tau_y]_2 (which = 1) as POINTS -- a ceremony SRS has no known tau: 192-byte ZCash
 * uncompressed G2 (x.c1, x.c0, y.c1, y.c0), checked on the curve and in the prime-order subgroup */
int zkp_srs_import_g2(zkp_ctx* ctx, int which, const uint8_t g2_192[192]);
int zkp_srs_export_g2(zkp_ctx* ctx, int which, uint8_t g2_192[192]);
int zkp_srs_export_scale_point(zkp_ctx* ctx, uint32_t row, uint8_t out48[48]);
/* after zkp_srs_set_shape(local row length) and the imports of the slices: the rows are point-range shard `shard` of a
 * domain of 2^log_domain points (what zkp_srs_generate_shard records) */
int zkp_srs_set_shard(zkp_ctx* ctx, uint32_t log_domain, uint32_t shard);
/* `prover setup --generate-setup` (reference tests/conftest.py:50-65): the bivariate monomial SRS [tau_x^j tau_y^i]_1,
 * row i / column j, from a trapdoor (tests, local networks) ... */
int zkp_srs_generate_monomial2(zkp_ctx* ctx, const uint8_t tau_x_be[32], const uint8_t tau_y_be[32], uint32_t log_n,
                               uint32_t log_machines);
/* ... and `--generate-precompute`: monomial -> Lagrange IN PLACE by inverse group FFTs along Y and X -- no trapdoor,
 * so it also derives the worker rows U[i][j] = [R_i(tau_y) L_j(tau_x)]_1 and the scale points [R_i(tau_y)]_1 from a
 * ceremony SRS.  Byte-identical to zkp_srs_generate with the same trapdoor. */
int zkp_srs_monomial_to_lagrange(zkp_ctx* ctx);
int zkp_srs_save(zkp_ctx* ctx, const char* path);
int zkp_srs_load(zkp_ctx* ctx, const char* path);
int zkp_srs_shape(zkp_ctx* ctx, uint32_t* log_n, uint32_t* log_machines);

/* ---- the hot path */
/* Client.worker_commit(i, poly)  (reference neurons/miner.py:38-45): poly = n evaluations, out = com_i */
int zkp_worker_commit(zkp_ctx* ctx, uint32_t i, const uint8_t* poly_be, size_t n, uint8_t commitment48[48]);
/* Client.worker_open(i, poly, x)  (reference neurons/miner.py:47-54): y = f_i(x), proof = [q_i]_1 */
int zkp_worker_open(zkp_ctx* ctx, uint32_t i, const uint8_t* poly_be, size_t n, const uint8_t x_be[32],
                    uint8_t eval_be[32], uint8_t proof48[48]);
/* The second half of the reference's two-call flow (neurons/miner.py:56-61: rpc_commit(i, poly), then
 * rpc_open(i, poly, alpha) with the SAME poly): opens the polynomial that the previous zkp_worker_commit /
 * zkp_worker_open / zkp_worker_commit_open call on this context left on the device, without a second upload.
 * ZKP_ERR_STATE when no polynomial of n elements is resident (any other call that stages scalars drops it). */
int zkp_worker_open_resident(zkp_ctx* ctx, uint32_t i, size_t n, const uint8_t x_be[32], uint8_t eval_be[32],
                             uint8_t proof48[48]);
/* zkp_worker_open_resident bound to ONE upload: zkp_resident_generation, called right after the caller's own
 * zkp_worker_commit / zkp_worker_open / zkp_worker_commit_open, names that upload; every later call that rewrites the
 * staged scalars (another client of a shared context, another thread, a raw zkp_msm_g1) changes the generation and
 * zkp_worker_open_resident_gen then fails with ZKP_ERR_STATE instead of opening somebody else's polynomial. */
int zkp_resident_generation(zkp_ctx* ctx, uint64_t* generation, size_t* n);
int zkp_worker_open_resident_gen(zkp_ctx* ctx, uint32_t i, size_t n, uint64_t generation, const uint8_t x_be[32],
                                 uint8_t eval_be[32], uint8_t proof48[48]);
/* Staged upload: the polynomial goes to the device in chunks while the host is still producing it (the shim decodes the
 * List[str] of the Prove synapse, reference base/protocol.py:35-40, chunk by chunk, and every finished chunk is on its way
 * over PCIe while the next is being decoded).  zkp_stage_begin names the upload; zkp_stage_chunk enqueues the asynchronous
 * copy of elements [first, first + count) taken from base + 32 first (page-locked memory; the argument order makes it the
 * per-chunk callback of the wire decoder); zkp_stage_end marks the polynomial resident.  zkp_worker_commit_resident,
 * zkp_worker_open_resident_gen and zkp_worker_commit_open_resident then work on that upload (or on the one an earlier
 * zkp_worker_* call left) and fail with ZKP_ERR_STATE if anything has replaced it. */
int zkp_stage_begin(zkp_ctx* ctx, size_t n, uint64_t* generation);
int zkp_stage_chunk(zkp_ctx* ctx, size_t first, const uint8_t* base, size_t count);
int zkp_stage_end(zkp_ctx* ctx, uint64_t generation);
int zkp_worker_commit_resident(zkp_ctx* ctx, uint32_t i, size_t n, uint64_t generation, uint8_t commitment48[48]);
int zkp_worker_commit_open_resident(zkp_ctx* ctx, uint32_t i, size_t n, uint64_t generation, const uint8_t x_be[32],
                                    uint8_t commitment48[48], uint8_t eval_be[32], uint8_t proof48[48]);
/* Miner.rpc_commit_and_open fused (reference neurons/miner.py:56-61): one upload, both MSMs */
int zkp_worker_commit_open(zkp_ctx* ctx, uint32_t i, const uint8_t* poly_be, size_t n, const uint8_t x_be[32],
                           uint8_t commitment48[48], uint8_t eval_be[32], uint8_t proof48[48]);
/* `count` independent commit+open requests of one row length (n = 2^log_n) in ONE launch set: the serving form of
 * Miner.forward for the live workload (2^16-element rows, reference Makefile:64-74; requests of several validators in
 * flight on the axon's threads, base/miner.py:66-70).  Request r: worker index rows[r], evaluations polys_be[r]
 * (n x 32 bytes, host), point xs_be + 32 r; outputs at index r.  status[r] is ZKP_OK or the error of request r alone
 * (ZKP_ERR_ENCODING for a non-canonical element: its outputs are zeroed, the others are unaffected); the return
 * value reports failures of the whole call.  Byte-identical to `count` calls of zkp_worker_commit_open. */
int zkp_worker_commit_open_batch(zkp_ctx* ctx, size_t count, const uint32_t* rows, const uint8_t* const* polys_be, size_t n,
                                 const uint8_t* xs_be, uint8_t* commitments48, uint8_t* evals_be, uint8_t* proofs48, int* status);
/* Client.worker_verify(i, proof, alpha, eval, commitment)  (reference neurons/validator.py:77-86).
 * Malformed / off-curve / wrong-subgroup points give *valid = 0 with status ZKP_OK
 * (reference tests/test_validator.py:66,79-86 expect reward 0, not an exception). */
int zkp_worker_verify(zkp_ctx* ctx, uint32_t i, const uint8_t proof48[48], const uint8_t alpha_be[32],
                      const uint8_t eval_be[32], const uint8_t commitment48[48], int* valid);
/* The responses of ONE challenge (common alpha) verified together: a random linear combination, two Miller loops
 * and one final exponentiation for the whole batch instead of per response (the reference scores responses one
 * by one, neurons/validator.py:168-170,178-192).  valid[k] is exactly what zkp_worker_verify would say for item k
 * (if the combined check fails the items are re-verified individually). */
int zkp_worker_verify_batch(zkp_ctx* ctx, size_t count, const uint32_t* indices, const uint8_t* proofs48,
                            const uint8_t alpha_be[32], const uint8_t* evals_be, const uint8_t* commitments48, int* valid);
/* Client.fft(poly, left, inverse)  (reference neurons/validator.py:58-65): natural-order (i)NTT over the
 * size-n X-domain (left != 0) or Y-domain (left == 0); n a power of two */
int zkp_fft(zkp_ctx* ctx, const uint8_t* in_be, size_t n, int left, int inverse, uint8_t* out_be);
/* Client.eval(poly, x)  (reference neurons/validator.py:97-104): coefficient-form Horner */
int zkp_eval(zkp_ctx* ctx, const uint8_t* coeffs_be, size_t n, const uint8_t x_be[32], uint8_t y_be[32]);
/* Validator.generate_challenge's evaluations in one call (reference neurons/validator.py:106-120: per row one
 * inverse fft and one eval through two RPCs): evals[i] = f_i(alpha) for `rows` rows of n evaluations each. */
int zkp_challenge_evals(zkp_ctx* ctx, const uint8_t* polys_be, size_t rows, size_t n, const uint8_t alpha_be[32],
                        uint8_t* evals_be);
/* Client.random_poly() / random_point()  (reference neurons/validator.py:68-75,88-95) */
int zkp_random_poly(zkp_ctx* ctx, uint64_t seed, uint8_t* out_be, size_t count);
int zkp_random_point(zkp_ctx* ctx, uint64_t seed, uint8_t out_be[32]);
/* elements [first, first + count) of the stream zkp_random_poly(seed, ...) yields (each GPU of a sharded job generates
 * its own slice of one global vector) */
int zkp_random_poly_range(zkp_ctx* ctx, uint64_t seed, uint64_t first, uint8_t* out_be, size_t count);

/* ---- Pianist master node (eprint 2023/1271 section 3): what the validator does with the M = 2^log_machines
 *      worker responses once "multi-miner proofs" land (reference neurons/validator.py:198 "not yet
 *      implemented", README.md:38).  f(X,Y) = sum_i R_i(Y) f_i(X);  com = sum_i com_i and pi_X = sum_i pi_i are
 *      zkp_g1_sum over the workers' 48-byte answers.  zkp_master_open_y opens g(Y) = f(alpha, Y), given by its
 *      evaluations y_i = f_i(alpha) (the workers' `eval`s, M x 32 bytes, worker order), at beta:
 *      z = g(beta), pi_Y = [(g(tau_y) - z)/(tau_y - beta)]_1 over the row scale points [R_i(tau_y)]_1.
 *      zkp_master_verify checks e(com - [z]_1, g2) == e(pi_X, [tau_x - alpha]_2) e(pi_Y, [tau_y - beta]_2);
 *      malformed points give *valid = 0 with ZKP_OK, as for zkp_worker_verify.  Host arithmetic (M <= 2^16). */
int zkp_master_open_y(zkp_ctx* ctx, const uint8_t* worker_evals_be, size_t m, const uint8_t beta_be[32],
                      uint8_t z_be[32], uint8_t proof_y48[48]);
int zkp_master_verify(zkp_ctx* ctx, const uint8_t commitment48[48], const uint8_t proof_x48[48],
                      const uint8_t proof_y48[48], const uint8_t alpha_be[32], const uint8_t beta_be[32],
                      const uint8_t z_be[32], int* valid);

/* ---- multi-GPU inside the library: one process, one context and one persistent host thread per device, no PyTorch, no
 *      NCCL, no collective inside a kernel (the path shards with no exchange on the inner loop).
 *      What crosses devices: a 32-byte partial sum and two Jacobian points per GPU, through pinned host memory, added on
 *      the calling thread and compressed once.  This is what a miner with several GPUs behind ONE Client
 *      (reference base/miner.py:73-84) runs; the per-miner split it reproduces is neurons/validator.py:41-42,212-222.
 *        ZKP_LAYOUT_ROWS         every device holds the whole SRS; sub-polynomial k runs on device k mod G (Pianist)
 *        ZKP_LAYOUT_POINT_RANGE  device g holds points [g n/G, (g+1) n/G) of every row; ONE polynomial is split
 *      (G = number of devices; the point-range layout uses the largest power of two of them). */
typedef struct zkp_mgpu zkp_mgpu;
enum { ZKP_LAYOUT_ROWS = 1, ZKP_LAYOUT_POINT_RANGE = 2 };
enum { ZKP_MGPU_RESIDENT = 1 }; /* flags: reuse the inputs uploaded by the previous identical call (device-resident timing) */
int zkp_mgpu_create(const int* devices /* NULL = 0..count-1 */, int count, zkp_mgpu** out);
void zkp_mgpu_destroy(zkp_mgpu* mg);
int zkp_mgpu_device_count(zkp_mgpu* mg);
/* the context of device k, borrowed (SRS import, zkp_worker_verify, zkp_master_*, tuning knobs); not to be used while a
 * zkp_mgpu_* call is running.  After filling the contexts by hand, zkp_mgpu_set_layout records what they hold. */
zkp_ctx* zkp_mgpu_ctx(zkp_mgpu* mg, int k);
int zkp_mgpu_set_layout(zkp_mgpu* mg, int layout, uint32_t log_n, uint32_t log_machines);
int zkp_mgpu_srs_generate(zkp_mgpu* mg, const uint8_t tau_x_be[32], const uint8_t tau_y_be[32], uint32_t log_n,
                          uint32_t log_machines, int layout);
int zkp_mgpu_prebuild_tables(zkp_mgpu* mg);
/* one G1 MSM of n <= 2^log_n points split by point range; scalars_be = the FULL vector on the host */
int zkp_mgpu_msm_g1(zkp_mgpu* mg, uint32_t row, const uint8_t* scalars_be, size_t n, int flags, uint8_t out48[48]);
/* worker_commit + worker_open of ONE polynomial of exactly 2^log_n evaluations, split by point range */
int zkp_mgpu_commit_open(zkp_mgpu* mg, uint32_t row, const uint8_t* poly_be, size_t n, const uint8_t x_be[32], int flags,
                         uint8_t commitment48[48], uint8_t eval_be[32], uint8_t proof48[48]);
/* Pianist: `count` sub-polynomials (polys_be: count x n x 32 bytes), sub-polynomial k = SRS row rows[k] on device
 * k mod G, opened at the common alpha; per-worker answers plus the master node's aggregates com = sum com_k,
 * pi_X = sum pi_k (either may be NULL).  zkp_master_open_y / zkp_master_verify complete the bivariate opening. */
int zkp_mgpu_pianist_commit_open(zkp_mgpu* mg, const uint32_t* rows, size_t count, const uint8_t* polys_be, size_t n,
                                 const uint8_t alpha_be[32], int flags, uint8_t* commitments48, uint8_t* evals_be,
                                 uint8_t* proofs48, uint8_t agg_commitment48[48], uint8_t agg_proof48[48]);

/* ---- wire codec for the List[str] format of the Prove synapse (reference base/protocol.py:35-40):
 *      `strs` holds `count` base64 strings of 43 (unpadded) or 44 (padded) characters each,
 *      concatenated with a fixed stride; output is count x 32 bytes.  Encoding emits 43 chars/elt. */
int zkp_b64_decode_fr(const char* strs, size_t stride, size_t count, uint8_t* out_be);
int zkp_b64_encode_fr(const uint8_t* in_be, size_t count, char* out_strs /* count*43 */);

/* ---- raw / benchmark entries (device-resident operands, CUDA-event timed inside the library) */
/* G1 MSM over the first n points of SRS row `row` with n scalars (big-endian) */
int zkp_msm_g1(zkp_ctx* ctx, uint32_t row, const uint8_t* scalars_be, size_t n, uint8_t out48[48]);
/* upload scalars once, then run `reps` MSMs back to back; returns mean device ms per MSM */
int zkp_bench_msm(zkp_ctx* ctx, uint32_t row, const uint8_t* scalars_be, size_t n, int reps, int flush_l2,
                  float* ms_per_msm, uint8_t out48[48]);
int zkp_bench_commit_open(zkp_ctx* ctx, uint32_t row, const uint8_t* poly_be, size_t n, const uint8_t x_be[32],
                          int reps, int flush_l2, float* ms_per_iter, float* ms_msm_kernel, uint32_t* launches,
                          uint8_t commitment48[48], uint8_t eval_be[32], uint8_t proof48[48]);
int zkp_bench_last_kernel_ms(zkp_ctx* ctx, float* ms);
/* flush the L2 of the context's device (256 MiB memset) and wait: for timing loops outside the library */
int zkp_bench_flush_l2(zkp_ctx* ctx);
/* stage timeline of one commit+open (CUDA events between the pipeline stages of both lanes): text lines
 * "<lane> <stage> <ms since request start>" and a final "host total <ms>" */
int zkp_bench_trace(zkp_ctx* ctx, uint32_t row, const uint8_t* poly_be, size_t n, const uint8_t x_be[32], int warm,
                    char* out, size_t out_cap);
int zkp_bench_ntt(zkp_ctx* ctx, size_t n, int reps, int inverse, float* ms_per_ntt);
/* roofline denominators measured live: chip-wide IMAD.WIDE.U32 issue rate and dependent-chain Fq products/s */
int zkp_bench_peaks(zkp_ctx* ctx, double* imad_wide_per_s, double* fq_mul_per_s);
/* MSM tuning knobs: window bits (0 = automatic) */
int zkp_set_msm_window(zkp_ctx* ctx, uint32_t c);
/* How the (bucket, point) entries of an MSM are grouped: 1 = this library's counting sort (histogram, scan, scatter;
 * the order inside a bucket is irrelevant), 0 = cub::DeviceRadixSort, 2 (default) = by size (the counting sort
 * while its scattered output stays cache-resident).  Results are identical for every setting. */
int zkp_set_msm_sort(zkp_ctx* ctx, int bucket_sort);
/* 1 (default): keep per-row fixed-base tables [2^(c w)] P_i in HBM (W x the row) so that all digit positions
 * share one bucket set; 0: classic per-window buckets.  Results are identical either way. */
int zkp_set_msm_mode(zkp_ctx* ctx, int fixed_base_tables);
/* rounds of batched-affine pairwise additions in front of the XYZZ bucket accumulation (6 instead of 10 field
 * products per addition, one shared inversion per 64 additions): -1 / 0 = off (default: measured no faster on
 * B200), 1..6 = that many rounds.  Results are identical for every setting. */
int zkp_set_msm_affine_rounds(zkp_ctx* ctx, int rounds);
int zkp_msm_info(zkp_ctx* ctx, size_t n, uint32_t* c, uint32_t* windows, uint64_t* fq_muls);
/* Convention switch (SURVEY.md section 8c "choose, document, keep switchable"): 0 (default) = the `poly` of
 * zkp_worker_commit / open / commit_open holds EVALUATIONS on the natural-order domain (implied by the reference's own
 * flow: the validator checks eval(fft(poly, inverse), alpha) against a proof made from the untransformed poly,
 * neurons/validator.py:116-117 vs :41-42); 1 = it holds COEFFICIENTS (the wording of the comment at
 * neurons/validator.py:67).  In coefficient form the library evaluates first (one forward NTT) and proceeds identically;
 * the polynomial must then fill the row exactly. */
int zkp_set_poly_form(zkp_ctx* ctx, int coefficients);
/* Experiment switch: the contiguous pass-2 tile of the NTT fetched by ONE bulk async copy (TMA: cp.async.bulk +
 * mbarrier) instead of per-thread 16-byte loads.  Identical results; off by default because it measured slower. */
int zkp_set_ntt_tma(zkp_ctx* ctx, int on);
/* commit+open as ONE grouped launch set (both MSMs share the sort, the accumulation grid, the slot levels and the
 * reduction): 1 = always, 0 = never (two streams, two launch sets), -1 (default) = by row length.  Same bytes out. */
int zkp_set_fuse(zkp_ctx* ctx, int mode);
/* Opening (the field half of zkp_worker_open, reference neurons/miner.py:47-54) of a SINGLE request: 1 (default) = the
 * blocks of pass 1 own cosets of the domain, whose products are closed forms x^m - w^(bm), and the one inversion
 * 1/(x^n - 1) is done on the host; 0 = contiguous runs with one Fermat inversion per block on the device (the form batches
 * and point-range shards always use).  Same bytes out; the first form removes ~0.2 ms of inversion latency per opening. */
int zkp_set_open_coset(zkp_ctx* ctx, int on);
/* Bucket reduction of an MSM over a SMALL bucket array (one request at the mainnet row size, 2^15 buckets): 1 (default) =
 * the row / column sums are formed by four lanes per share and two warps per sum (both stages are pure latency there);
 * 0 = the kernels used for large arrays (one thread per share, one warp per sum).  Same bytes out. */
int zkp_set_rowcol_coop(zkp_ctx* ctx, int on);
/* Fixed-base tables live in one arena of equal slots (one per row, as many as fit the budget; least-recently-used
 * rows are evicted when the arena is smaller than the SRS).  zkp_srs_prebuild_tables builds the tables of rows
 * [first_row, first_row + count) now rather than inside the first request that needs them (*built = tables resident
 * afterwards); zkp_set_table_budget caps the arena in bytes (0 = 60% of the HBM free at first use; the environment
 * variable ZKP_B200_TABLE_BYTES overrides both); zkp_srs_table_stats reports {resident tables, slots, arena bytes,
 * builds, evictions, calls that fell back to the classic per-window path}.  A fallback is logged to stderr once per
 * SRS, never silent. */
int zkp_srs_prebuild_tables(zkp_ctx* ctx, uint32_t first_row, uint32_t count, uint32_t* built);
int zkp_set_table_budget(zkp_ctx* ctx, size_t bytes);
int zkp_srs_table_stats(zkp_ctx* ctx, uint64_t out[6]);

/* ---- pairing check exposed for tests: prod e(P_k, Q_k) == 1, P compressed G1 (48 B), Q affine G2 as
 *      4 x 48 B big-endian (x.c0, x.c1, y.c0, y.c1); G2 inputs are checked on-curve and in the prime-order subgroup */
int zkp_pairing_check(const uint8_t* g1_48, const uint8_t* g2_192, size_t pairs, int* is_one);

#ifdef __cplusplus
}
#endif
#endif /* ZKP_B200_H */
