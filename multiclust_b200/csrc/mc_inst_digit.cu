/* mc_inst_digit.cu -- instantiations of digit_kernel<K, MODE> */
#include "mc_dispatch.h"

template <int MODE> static digit_fn pick_k(int K)
{
	switch (K) {
	case 1: return digit_kernel<1, MODE>;
	case 2: return digit_kernel<2, MODE>;
	case 3: return digit_kernel<3, MODE>;
	case 4: return digit_kernel<4, MODE>;
	case 5: return digit_kernel<5, MODE>;
	case 6: return digit_kernel<6, MODE>;
	case 7: return digit_kernel<7, MODE>;
	case 8: return digit_kernel<8, MODE>;
	case 9: return digit_kernel<9, MODE>;
	case 10: return digit_kernel<10, MODE>;
	case 11: return digit_kernel<11, MODE>;
	case 12: return digit_kernel<12, MODE>;
	case 13: return digit_kernel<13, MODE>;
	case 14: return digit_kernel<14, MODE>;
	case 15: return digit_kernel<15, MODE>;
	case 16: return digit_kernel<16, MODE>;
	}
	return nullptr;
}

digit_fn mc_pick_digit(int K, int mode)
{
	return mode == DG_MIX_E ? pick_k<DG_MIX_E>(K) : mode == DG_MIX_M ? pick_k<DG_MIX_M>(K) : nullptr;
}
