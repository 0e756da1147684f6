// remaining C-ABI entries (filled in as the path widens)
#pragma once
extern "C" {
#define ZKP_TODO(name, ...) int name(__VA_ARGS__) { return zkp::fail(ZKP_ERR_STATE, #name ": not implemented yet"); }
ZKP_TODO(zkp_srs_generate, zkp_ctx*, const uint8_t*, const uint8_t*, uint32_t, uint32_t)
ZKP_TODO(zkp_srs_save, zkp_ctx*, const char*)
ZKP_TODO(zkp_srs_load, zkp_ctx*, const char*)
ZKP_TODO(zkp_worker_open, zkp_ctx*, uint32_t, const uint8_t*, size_t, const uint8_t*, uint8_t*, uint8_t*)
ZKP_TODO(zkp_worker_commit_open, zkp_ctx*, uint32_t, const uint8_t*, size_t, const uint8_t*, uint8_t*, uint8_t*, uint8_t*)
ZKP_TODO(zkp_worker_verify, zkp_ctx*, uint32_t, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, int*)
ZKP_TODO(zkp_fft, zkp_ctx*, const uint8_t*, size_t, int, int, uint8_t*)
ZKP_TODO(zkp_eval, zkp_ctx*, const uint8_t*, size_t, const uint8_t*, uint8_t*)
ZKP_TODO(zkp_random_poly, zkp_ctx*, uint64_t, uint8_t*, size_t)
ZKP_TODO(zkp_random_point, zkp_ctx*, uint64_t, uint8_t*)
ZKP_TODO(zkp_b64_decode_fr, const char*, size_t, size_t, uint8_t*)
ZKP_TODO(zkp_b64_encode_fr, const uint8_t*, size_t, char*)
ZKP_TODO(zkp_bench_msm, zkp_ctx*, uint32_t, const uint8_t*, size_t, int, int, float*, uint8_t*)
ZKP_TODO(zkp_bench_commit_open, zkp_ctx*, uint32_t, const uint8_t*, size_t, const uint8_t*, int, int, float*, float*, uint32_t*, uint8_t*, uint8_t*, uint8_t*)
ZKP_TODO(zkp_bench_ntt, zkp_ctx*, size_t, int, int, float*)
ZKP_TODO(zkp_pairing_check, const uint8_t*, const uint8_t*, size_t, int*)
}
