// u8_scan_tc.cuh -- tcgen05 kind::i8 exact L2 scan over uint8 rows (definitions in u8_scan_tc.cu).
#pragma once
#include "common.cuh"

namespace b200nn {

bool u8_scan_tc_supported(int D, int k);
// rows [n][D] u8 -> canonical 256-row tiles + per-tile row meta xmeta[tiles][2][256] = {|x|^2, label rank}
// (n_pad = multiple of 256; xmeta holds 2 * n_pad ints)
int launch_u8_rows_to_canonical(Ctx* ctx, const unsigned char* rows, const uint32_t* rank, long long n, int D, unsigned char* xcan,
                                int* xmeta, long long n_pad);
// row slices for a pass over n_tiles tiles of 256 rows (one wave of CTAs, at least min_tiles_per_slice tiles each)
int u8_scan_tc_slices(int sm_count, long long nq, long long n_tiles, int min_tiles_per_slice = 4);
int u8_scan_tc_lists_per_slice(int D, int k);
// out_keys [n_slices * u8_scan_tc_lists_per_slice(D, k)][nq][k]; ids in the keys are label ranks.
// init_thr (may be NULL): init_thr[q * init_stride] = an upper bound on query q's k-th best distance, e.g. the k-th
// best over a prefix of the rows -- rows beyond it are dropped (ties kept), so lists may come back shorter than k.
// The launch scans tiles [tile0, tile0 + n_tiles) of the index (n = rows indexed, for the validity of the last tile).
// gmin (may be NULL): [nq][u8_scan_tc_bound_lists()] ints preset to 0x7f7f7f7f -- one-pass shared-bound mode: the first
// lists of every query publish their minimum there and a bound warp per CTA turns them into a running upper bound on the
// k-th best distance (k-th smallest of the minima); needs k <= min(bound lists, n_slices * lists per slice).
int u8_scan_tc_bound_lists();
int launch_u8_scan_tc(Ctx* ctx, const unsigned char* xcan, const int* xmeta, long long n, long long tile0, long long n_tiles, int D,
                      const unsigned char* queries, long long nq, int n_slices, int k, const int* init_thr, int init_stride, int* gmin,
                      unsigned long long* out_keys);

}  // namespace b200nn
