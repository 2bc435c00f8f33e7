/* mc_inst_dense.cu -- instantiations of dense_kernel<NB, PMAX, MODE> and the
 * look-up of the admix3 modes */
#include "mc_dispatch.h"

template <int NB, int MODE> static dense_fn pick_pmax(int pmax)
{
	switch (pmax) {
	case 1: return dense_kernel<NB, 1, MODE>;
	case 2: return dense_kernel<NB, 2, MODE>;
	case 4: return dense_kernel<NB, 4, MODE>;
	case 7: return dense_kernel<NB, 7, MODE>;
	case 15: return dense_kernel<NB, 15, MODE>;
	}
	return nullptr;
}

dense_fn mc_pick_dense(int NB, int pmax, int mode)
{
	switch (mode) {
	case DN_ADMIX_EM:
		return NB == 1 ? pick_pmax<1, DN_ADMIX_EM>(pmax) : pick_pmax<2, DN_ADMIX_EM>(pmax);
	case DN_ADMIX_LL:
		return NB == 1 ? pick_pmax<1, DN_ADMIX_LL>(pmax) : pick_pmax<2, DN_ADMIX_LL>(pmax);
	case DN_MIX_E:
		return NB == 1 ? dense_kernel<1, 1, DN_MIX_E> : dense_kernel<2, 1, DN_MIX_E>;
	case DN_MIX_M:
		return NB == 1 ? dense_kernel<1, 1, DN_MIX_M> : dense_kernel<2, 1, DN_MIX_M>;
	}
	return nullptr;
}

admix3_fn mc_pick_admix3(int mode, int KP, int PP)
{
	switch (mode) {
	case A3_ADMIX_EM: return mc_pick_admix3_em(KP, PP);
	case A3_ADMIX_LL: return mc_pick_admix3_ll(KP, PP);
	case A3_MIX_E: return mc_pick_admix3_mix_e(KP, PP);
	case A3_MIX_M: return mc_pick_admix3_mix_m(KP, PP);
	}
	return nullptr;
}
