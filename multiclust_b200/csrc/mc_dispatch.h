/*
 * mc_dispatch.h -- the genotype-streaming kernels are instantiated in their own
 * translation units (mc_inst_*.cu, compiled in parallel); mc_cuda.cu reaches
 * them through these look-ups.  A null result means "no such instantiation".
 */
#pragma once

#include "mc_admix3.cuh"
#include "mc_dense.cuh"
#include "mc_digit.cuh"
#include "mc_tile.cuh"

typedef void (*admix3_fn)(const Admix3Args);
typedef void (*dense_fn)(const DenseArgs);
typedef void (*tile_fn)(const TileArgs);
typedef void (*digit_fn)(const DigitArgs);

/* mode: A3_ADMIX_EM, A3_ADMIX_LL, A3_MIX_E, A3_MIX_M; KP = ceil(K / 2) in 1..8;
 * PP = padded ploidy 1, 2, 4, 8 */
admix3_fn mc_pick_admix3(int mode, int KP, int PP);
admix3_fn mc_pick_admix3_em(int KP, int PP);
admix3_fn mc_pick_admix3_ll(int KP, int PP);
admix3_fn mc_pick_admix3_mix_e(int KP, int PP);
admix3_fn mc_pick_admix3_mix_m(int KP, int PP);
/* mode: DN_*; NB = 1, 2; pmax = 1, 2, 4, 7, 15 (largest allele count) */
dense_fn mc_pick_dense(int NB, int pmax, int mode);
/* mode: MODE_*; KH = 1..6; PP = 1, 2, 4, 8, 16 */
tile_fn mc_pick_tile(int mode, int KH, int PP);
/* mode: DG_MIX_E, DG_MIX_M; K = 1..16 */
digit_fn mc_pick_digit(int K, int mode);
