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
/* the same with 96-byte ZCash-uncompressed inputs (no square roots on the combining rank); zkp_g1_uncompress
 * expands a compressed point on the rank that produced it */
int zkp_g1_uncompress(const uint8_t in48[48], uint8_t out96[96]);
int zkp_g1_sum_uncompressed(const uint8_t* points96, size_t count, uint8_t out48[48]);
/* import one row from 96-byte ZCash-uncompressed points (validated on curve), and its scale point */
int zkp_srs_set_shape(zkp_ctx* ctx, uint32_t log_n, uint32_t log_machines);
int zkp_srs_import_row(zkp_ctx* ctx, uint32_t row, const uint8_t* points96, size_t n,
                       const uint8_t scale_point48[48]);
int zkp_srs_import_g2_tau(zkp_ctx* ctx, const uint8_t tau_x_be[32]);
int zkp_srs_import_g2_tau_y(zkp_ctx* ctx, const uint8_t tau_y_be[32]); /* [tau_y]_2, master verification only */
int zkp_srs_export_row(zkp_ctx* ctx, uint32_t row, uint8_t* points96, size_t n);
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
/* Miner.rpc_commit_and_open fused (reference neurons/miner.py:56-61): one upload, both MSMs */
int zkp_worker_commit_open(zkp_ctx* ctx, uint32_t i, const uint8_t* poly_be, size_t n, const uint8_t x_be[32],
                           uint8_t commitment48[48], uint8_t eval_be[32], uint8_t proof48[48]);
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

/* ---- pairing check exposed for tests: prod e(P_k, Q_k) == 1, P compressed G1 (48 B), Q affine G2 as
 *      4 x 48 B big-endian (x.c0, x.c1, y.c0, y.c1) */
int zkp_pairing_check(const uint8_t* g1_48, const uint8_t* g2_192, size_t pairs, int* is_one);

#ifdef __cplusplus
}
#endif
#endif /* ZKP_B200_H */
