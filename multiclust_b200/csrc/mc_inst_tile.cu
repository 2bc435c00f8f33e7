/* mc_inst_tile.cu -- instantiations of tile_kernel<KH, PP, MODE> */
#include "mc_dispatch.h"

template <int KH, int MODE> static tile_fn pick_pp(int PP)
{
	switch (PP) {
	case 1: return tile_kernel<KH, 1, MODE>;
	case 2: return tile_kernel<KH, 2, MODE>;
	case 4: return tile_kernel<KH, 4, MODE>;
	case 8: return tile_kernel<KH, 8, MODE>;
	case 16: return tile_kernel<KH, 16, MODE>;
	}
	return nullptr;
}

template <int MODE> static tile_fn pick_kh(int KH, int PP)
{
	switch (KH) {
	case 1: return pick_pp<1, MODE>(PP);
	case 2: return pick_pp<2, MODE>(PP);
	case 3: return pick_pp<3, MODE>(PP);
	case 4: return pick_pp<4, MODE>(PP);
	case 5: return pick_pp<5, MODE>(PP);
	case 6: return pick_pp<6, MODE>(PP);
	}
	return nullptr;
}

tile_fn mc_pick_tile(int mode, int KH, int PP)
{
	switch (mode) {
	case MODE_ADMIX_EM: return pick_kh<MODE_ADMIX_EM>(KH, PP);
	case MODE_ADMIX_LL: return pick_kh<MODE_ADMIX_LL>(KH, PP);
	case MODE_MIX_E: return pick_kh<MODE_MIX_E>(KH, PP);
	case MODE_MIX_M: return pick_kh<MODE_MIX_M>(KH, PP);
	}
	return nullptr;
}
