/* mc_inst_a3_em.cu -- instantiations of admix3_kernel<KP, PP, A3_ADMIX_EM> */
#include "mc_dispatch.h"

template <int KP> static admix3_fn pick_pp(int PP)
{
	switch (PP) {
	case 1: return admix3_kernel<KP, 1, A3_ADMIX_EM>;
	case 2: return admix3_kernel<KP, 2, A3_ADMIX_EM>;
	case 4: return admix3_kernel<KP, 4, A3_ADMIX_EM>;
	case 8: return admix3_kernel<KP, 8, A3_ADMIX_EM>;
	}
	return nullptr;
}

admix3_fn mc_pick_admix3_em(int KP, int PP)
{
	switch (KP) {
	case 1: return pick_pp<1>(PP);
	case 2: return pick_pp<2>(PP);
	case 3: return pick_pp<3>(PP);
	case 4: return pick_pp<4>(PP);
	case 5: return pick_pp<5>(PP);
	case 6: return pick_pp<6>(PP);
	case 7: return pick_pp<7>(PP);
	case 8: return pick_pp<8>(PP);
	}
	return nullptr;
}
