// In-tree device-wide primitives for the candidate lists of greedy_nms.cu and explore.cu (rows a21 / f3): a STABLE
// descending radix sort of 64-bit keys (optionally with 32-bit values) and an order-preserving compaction by byte flags.
// (Round 1 used cub::DeviceRadixSort / cub::DeviceSelect here.)
#pragma once
#include "common.cuh"

namespace cetpick {

// scratch bytes for radix_sort_desc_u64 / compact_flagged on up to n elements
size_t sort_tmp_bytes(size_t n, bool with_values);
size_t compact_tmp_bytes(size_t n);

// keys_out <- keys_in sorted descending, ties in input order; vals (nullable pair) travel with their keys.
// keys_in is not modified; keys_out / vals_out must not alias the inputs.  `launches` (nullable) += kernels enqueued.
int radix_sort_desc_u64(const unsigned long long* keys_in, unsigned long long* keys_out, const uint32_t* vals_in,
                        uint32_t* vals_out, uint32_t n, void* tmp, size_t tmp_bytes, cudaStream_t s, int64_t* launches);

// out[j] = src ? src[i] : i  for the j-th i with flags[i] != 0 (input order kept); *n_out (device) = number selected.
// Exactly one of out_u64 (gathers src) / out_idx (writes the index) is used.
int compact_flagged(const uint8_t* flags, uint32_t n, const unsigned long long* src, unsigned long long* out_u64,
                    uint32_t* out_idx, int* n_out, void* tmp, size_t tmp_bytes, cudaStream_t s, int64_t* launches);

}  // namespace cetpick
