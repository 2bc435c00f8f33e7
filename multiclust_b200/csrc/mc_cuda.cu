/*
 * mc_cuda.cu -- context, work planning and the C ABI of include/mc_cuda.h.
 *
 * Build (see multiclust_b200/build.py):
 *   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared \
 *        -Xcompiler -fPIC -Iinclude -o libmc_cuda.so mc_cuda.cu
 */
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include "mc_cuda.h"
#include "mc_dispatch.h"
#include "mc_kernels.cuh"
#include "mc_admix3_build.cuh"
#include "mc_dense_build.cuh"
#include "mc_digit_build.cuh"

#define KH_MAX 6
#define SMEM_LIMIT (227 * 1024)

static std::string g_create_error;

struct mc_ctx {
	int device = 0;
	cudaStream_t own_stream = nullptr, stream = nullptr;
	/* the eta side of an admixture step runs on a second stream, concurrently
	 * with the caller's exchange of the allele sums (mc_em_step_local) */
	cudaStream_t aux = nullptr;
	cudaEvent_t ev_main = nullptr, ev_aux = nullptr;
	bool aux_pending = false;
	bool fused_step = false;	/* mc_em_step: local + finish without an exchange */
	std::string err;
	int64_t launches = 0;
	int num_sms = 148;

	/* data */
	int64_t I = 0, T = 0;
	int L = 0, P = 0, PP = 0, IB = 0;
	std::vector<int32_t> J, off;
	unsigned char *d_nat = nullptr;		/* [I][L][P] */
	/* parametric bootstrap: the original data while d_nat holds a bootstrap
	 * sample, and the estimates under H0 the samples are drawn from */
	unsigned char *d_nat_orig = nullptr;
	double *d_mle_eta = nullptr, *d_mle_p = nullptr;
	int mle_K = 0, mle_admixture = 0, mle_per_indiv = 0;
	int *d_J = nullptr, *d_off = nullptr;

	/* model */
	bool have_model = false;
	int K = 0, admixture = 0, eta_constrained = 0, per_indiv = 0, q = 0,
		do_proj = 1;
	double eta_lb = 0, p_lb = 0;
	int64_t neta = 0, np = 0;
	double *d_p[3] = { nullptr, nullptr, nullptr };
	double *d_eta[3] = { nullptr, nullptr, nullptr };
	std::vector<double *> d_up, d_vp, d_ue, d_ve;
	double *d_post = nullptr;	/* D_ik or v_ik [I][K] */
	double *d_lli = nullptr;	/* mixture: per-block partial sums of the tail */
	long long lli_n = 0;
	double *d_logp = nullptr;	/* mixture: log p table [K][T] */

	/* plan */
	TileArgs ta;
	int k_split = 1, KH = 1, grid = 0, block = 0;
	size_t smem_em = 0;
	int *d_slot_locus = nullptr, *d_slot_off = nullptr, *d_slot_J = nullptr,
		*d_group_rowbase = nullptr, *d_group_rows = nullptr,
		*d_tile_rows = nullptr;
	unsigned char *d_tiled = nullptr;
	double *d_Apart = nullptr, *d_Npart = nullptr, *d_llpart = nullptr;
	double *d_xbuf = nullptr;	/* [K*T | ll | K] */
	double *d_red = nullptr;	/* reduction partials */
	double *d_small = nullptr;	/* small results for the host */
	int *d_IK = nullptr;

	/* two-pass admixture plan (mc_admix3.cuh); used when `use3` */
	bool use3 = false;
	bool layout3 = false;		/* codes / lists / column tables are built */
	int l3_ncolmax = 0, l3_max_tile_rows = 0;
	Admix3Args a3;
	int KP3 = 0, grid3 = 0;
	size_t smem3 = 0, smem3_ll = 0;
	int *d3_lt_ncol = nullptr, *d3_lc_first = nullptr;
	double *d3_Gacc = nullptr;
	unsigned short *d3_colinfo = nullptr, *d3_csc = nullptr, *d3_colstart = nullptr;
	unsigned char *d3_codes = nullptr;
	unsigned char *d3_perm_of = nullptr;	/* [T] allele slot -> row of its locus in the kernel */
	int *d3_nat_of = nullptr;		/* [T] kernel row -> allele slot (both from the start of p) */
	bool l3_permuted = false;		/* some locus has its rows reordered */
	int l3_cap = 0;				/* 16-bit entries per tile of the pass-2 lists */
	double *d3_pperm = nullptr;		/* [K][T] the parameter slot in kernel row order */
	/* dense DMMA plan for biallelic data (mc_dense.cuh); used when `use_dn` */
	bool use_dn = false;
	bool layout_dn = false;		/* packed counts are built */
	int dn_maxcode = -1;		/* largest allele code in the data, -1: not looked at yet */
	DenseArgs dn;
	int dn_NB = 1, dn_pbits = 1, grid_dn = 0;
	unsigned char *d_dn_cnt = nullptr;
	double *d_dn_pd = nullptr;
	int *d_dn_lc_first = nullptr;
	/* digit-sliced integer plan of the mixture model on the same data
	 * (mc_digit.cuh); used when `use_dg`, the dense plan is its fall-back */
	bool use_dg = false;
	bool layout_dg = false;		/* both count layouts are built */
	DigitArgs dgE, dgM;
	uint4 *d_dg_cntE = nullptr, *d_dg_cntM = nullptr;
	uint2 *d_dg_tabE = nullptr, *d_dg_tabM = nullptr;
	int *d_dg_flag = nullptr;	/* [0] table not representable, [1] chunks of the E pass */
	double *d_dg_vscale = nullptr;	/* [2 K] posterior column scaling, k_mix_final */
	int dn_nl = 0, dn_ni = 0;	/* chunk counts of the dense plan */
	int dg_kind = 0;		/* layout built: 1 biallelic (locus pairs), 2 column pairs */
	int dg_min_tiles = 0, dg_min_chunks = 0;	/* partial-sum slots the digit plan needs */
	int *d_col_locus = nullptr;	/* [T] locus of every allele column (column-pair form) */
	/* kernels whose dynamic shared-memory limit has been raised (set once per
	 * kernel and size, not on every launch) */
	std::vector<std::pair<const void *, size_t>> smem_attr;
	/* whole steps replayed as CUDA graphs (mc_em_step / mc_loglik on one
	 * context): a small fit is launch-bound, and a graph launch costs one
	 * submission instead of five to twelve */
	cudaGraphExec_t g_step[9] = {}, g_ll[3] = {};
	int64_t g_step_n[9] = {}, g_ll_n[3] = {};	/* kernels inside each graph */
	int g_step_seen[9] = {}, g_ll_seen[3] = {};
	bool graphs_ok = true;
	int opt_graph = 1;
	/* plan options (mc_set_option) */
	int opt_kernel = 0, opt_timing = 0;
	/* scratch of the admixture initialiser, kept between fits */
	unsigned char *d_init_z = nullptr;
	unsigned *d_init_N = nullptr, *d_init_h = nullptr;
	size_t init_z_n = 0, init_N_n = 0, init_h_n = 0;
	/* sizes of the partial-sum buffers of the active plan */
	int act_tiles = 0, act_chunks = 0, act_units = 0;
	long long act_Ipad = 0;

	/* profiling of the streaming kernel */
	bool profile = false;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
	int64_t prof_n = 0;
	double prof_ms = 0;
};

/* NVTX range around an ABI call (visible in Nsight Systems / ncu --nvtx;
 * a few nanoseconds when no tool is attached) */
struct NvtxRange {
	explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
	~NvtxRange() { nvtxRangePop(); }
};
#define NVTX_FN() NvtxRange nvtx_range_(__func__)

/* ---------------------------------------------------------------- errors */

static int fail(mc_ctx *c, int code, const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	if (c)
		c->err = buf;
	else
		g_create_error = buf;
	return code;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return fail(c, MC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
		cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

#define LAUNCH_CHECK(name) do { cudaError_t e_ = cudaGetLastError(); \
	c->launches++; if (e_ != cudaSuccess) \
	return fail(c, MC_ERR_CUDA, "launch of %s failed: %s", name, \
		cudaGetErrorString(e_)); } while (0)

extern "C" const char *mc_last_error(const mc_ctx *c)
{
	return c ? c->err.c_str() : g_create_error.c_str();
}

extern "C" int mc_abi_version(void) { return MC_ABI_VERSION; }

static int grid_for(const mc_ctx *c, long long n, int threads)
{
	long long g = (n + threads - 1) / threads;
	long long cap = (long long)c->num_sms * 8;
	if (g > cap) g = cap;
	if (g < 1) g = 1;
	return (int)g;
}

/* Device memory comes from the device's stream-ordered pool with the release
 * threshold lifted (mc_create): what a model or plan frees stays mapped and is
 * handed out again by the next allocation.  A K sweep re-allocates ~15 buffers
 * per K, and cudaMalloc / cudaFree cost 5-15 ms each on this platform. */
template <typename T> static cudaError_t mc_dev_malloc(cudaStream_t stream, T **p, size_t n)
{
	cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(p), n ? n : 1, stream);
	if (e != cudaSuccess) {	/* pool unsupported or exhausted: plain allocation */
		(void)cudaGetLastError();
		e = cudaMalloc(reinterpret_cast<void **>(p), n ? n : 1);
	}
	return e;
}
#define MC_DEV_MALLOC(ptr, n) mc_dev_malloc(c->stream, (ptr), (n))

template <typename T> static void dfree(T *&p)
{
	if (p)
		cudaFree(p);
	p = nullptr;
}

/* --------------------------------------------------------------- context */

extern "C" int mc_create(mc_ctx **out, int device)
{
	mc_ctx *c = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0)
		return fail(nullptr, MC_ERR_CUDA, "no CUDA device: %s (the MULTICLUST "
			"EM path has no CPU fallback)", cudaGetErrorString(e));
	if (device < 0 || device >= n)
		return fail(nullptr, MC_ERR_ARG, "device %d out of range (0..%d)",
			device, n - 1);
	c = new mc_ctx();
	c->device = device;
	if ((e = cudaSetDevice(device)) != cudaSuccess
		|| (e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
		delete c;
		return fail(nullptr, MC_ERR_CUDA, "cannot initialise device %d: %s",
			device, cudaGetErrorString(e));
	}
	c->stream = c->own_stream;
	if (cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking) != cudaSuccess
		|| cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming) != cudaSuccess
		|| cudaEventCreateWithFlags(&c->ev_aux, cudaEventDisableTiming) != cudaSuccess) {
		delete c;
		return fail(nullptr, MC_ERR_CUDA, "cannot create the auxiliary stream");
	}
	cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);
	{	/* keep freed device memory in the pool (see mc_dev_malloc) */
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
			unsigned long long keep = ~0ULL;
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		}
		(void)cudaGetLastError();
	}
	if (MC_DEV_MALLOC(&c->d_small, 64 * sizeof(double)) != cudaSuccess) {
		delete c;
		return fail(nullptr, MC_ERR_NOMEM, "cudaMalloc failed");
	}
	*out = c;
	return MC_OK;
}

static void free_graphs(mc_ctx *c)
{
	for (int x = 0; x < 9; x++) {
		if (c->g_step[x])
			cudaGraphExecDestroy(c->g_step[x]);
		c->g_step[x] = nullptr;
		c->g_step_seen[x] = 0;
	}
	for (int x = 0; x < 3; x++) {
		if (c->g_ll[x])
			cudaGraphExecDestroy(c->g_ll[x]);
		c->g_ll[x] = nullptr;
		c->g_ll_seen[x] = 0;
	}
	c->graphs_ok = true;
}

static void free_plan(mc_ctx *c)
{
	free_graphs(c);
	dfree(c->d_slot_locus); dfree(c->d_slot_off); dfree(c->d_slot_J);
	dfree(c->d_group_rowbase); dfree(c->d_group_rows); dfree(c->d_tile_rows);
	dfree(c->d_tiled); dfree(c->d_Apart); dfree(c->d_Npart);
	dfree(c->d_llpart); dfree(c->d_xbuf); dfree(c->d_red);
	dfree(c->d3_lc_first); dfree(c->d3_Gacc); dfree(c->d3_pperm);
	dfree(c->d_dn_pd); dfree(c->d_dn_lc_first);
	dfree(c->d_dg_tabE); dfree(c->d_dg_tabM); dfree(c->d_dg_flag); dfree(c->d_dg_vscale);
	c->use3 = false;
	c->use_dn = false;
	c->use_dg = false;
	c->dg_min_tiles = c->dg_min_chunks = 0;
}

/* the admixture kernel's tile codes and entry lists depend on the data only,
 * not on K: they survive mc_alloc_model and are rebuilt by mc_set_data */
static void free_layout3(mc_ctx *c)
{
	dfree(c->d3_lt_ncol); dfree(c->d3_colinfo);
	dfree(c->d3_csc); dfree(c->d3_colstart);
	dfree(c->d3_codes); dfree(c->d3_perm_of); dfree(c->d3_nat_of);
	c->layout3 = false;
	c->l3_permuted = false;
	dfree(c->d_dn_cnt);
	c->layout_dn = false;
}

/* the count layouts of the digit-sliced mixture kernels: data only, too */
static void free_layout_dg(mc_ctx *c)
{
	dfree(c->d_dg_cntE); dfree(c->d_dg_cntM); dfree(c->d_col_locus);
	c->layout_dg = false;
	c->dg_kind = 0;
}

static void free_model(mc_ctx *c)
{
	for (int s = 0; s < 3; s++) {
		dfree(c->d_p[s]);
		dfree(c->d_eta[s]);
	}
	for (auto *x : c->d_up) cudaFree(x);
	for (auto *x : c->d_vp) cudaFree(x);
	for (auto *x : c->d_ue) cudaFree(x);
	for (auto *x : c->d_ve) cudaFree(x);
	c->d_up.clear(); c->d_vp.clear(); c->d_ue.clear(); c->d_ve.clear();
	dfree(c->d_post); dfree(c->d_lli); dfree(c->d_logp); dfree(c->d_IK);
	free_plan(c);
	c->have_model = false;
}

static void free_data(mc_ctx *c)
{
	free_model(c);
	free_layout3(c);
	free_layout_dg(c);
	dfree(c->d_init_z); dfree(c->d_init_N); dfree(c->d_init_h);
	c->init_z_n = c->init_N_n = c->init_h_n = 0;
	dfree(c->d_nat); dfree(c->d_J); dfree(c->d_off);
	dfree(c->d_nat_orig); dfree(c->d_mle_eta); dfree(c->d_mle_p);
	c->mle_K = 0;
	c->I = 0;
	c->dn_maxcode = -1;
}

extern "C" void mc_destroy(mc_ctx *c)
{
	if (!c)
		return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	free_data(c);
	dfree(c->d_small);
	for (auto &ev : c->prof_events) {
		cudaEventDestroy(ev.first);
		cudaEventDestroy(ev.second);
	}
	cudaStreamSynchronize(c->aux);
	cudaEventDestroy(c->ev_main);
	cudaEventDestroy(c->ev_aux);
	cudaStreamDestroy(c->aux);
	cudaStreamDestroy(c->own_stream);
	delete c;
}

extern "C" int mc_set_stream(mc_ctx *c, void *s)
{
	if (!c)
		return MC_ERR_ARG;
	CK(cudaStreamSynchronize(c->stream));
	c->stream = s ? (cudaStream_t)s : c->own_stream;
	return MC_OK;
}

extern "C" int mc_ctx_device(const mc_ctx *c) { return c ? c->device : -1; }
extern "C" void *mc_ctx_stream(const mc_ctx *c) { return c ? (void *)c->stream : nullptr; }

extern "C" int mc_sync(mc_ctx *c)
{
	if (!c)
		return MC_ERR_ARG;
	CK(cudaSetDevice(c->device));
	CK(cudaStreamSynchronize(c->stream));
	return MC_OK;
}

/* ------------------------------------------------------------------ data */

static int set_dims(mc_ctx *c, int64_t I, int32_t L, int32_t P, const int32_t *J)
{
	if (I <= 0 || L <= 0 || P <= 0)
		return fail(c, MC_ERR_ARG, "bad dimensions I=%lld L=%d P=%d",
			(long long)I, L, P);
	if (P > 16)
		return fail(c, MC_ERR_UNSUPPORTED, "ploidy %d > 16 is not supported", P);
	free_data(c);
	c->I = I; c->L = L; c->P = P;
	c->PP = 1;
	while (c->PP < P)
		c->PP <<= 1;
	c->IB = c->PP >= 2 ? 16 / c->PP : 8;
	c->J.assign(J, J + L);
	c->off.assign((size_t)L + 1, 0);
	for (int l = 0; l < L; l++) {
		if (J[l] < 0 || J[l] > 255)
			return fail(c, MC_ERR_UNSUPPORTED, "locus %d has %d allele "
				"slots; 8-bit codes allow at most 255", l, J[l]);
		c->off[l + 1] = c->off[l] + J[l];
	}
	c->T = c->off[L];
	CK(MC_DEV_MALLOC(&c->d_J, sizeof(int) * (size_t)L));
	CK(MC_DEV_MALLOC(&c->d_off, sizeof(int) * ((size_t)L + 1)));
	CK(cudaMemcpyAsync(c->d_J, c->J.data(), sizeof(int) * (size_t)L,
		cudaMemcpyHostToDevice, c->stream));
	CK(cudaMemcpyAsync(c->d_off, c->off.data(), sizeof(int) * ((size_t)L + 1),
		cudaMemcpyHostToDevice, c->stream));
	return MC_OK;
}

extern "C" int mc_set_data(mc_ctx *c, int64_t I, int32_t L, int32_t P,
	const int32_t *J, const uint8_t *codes)
{
	NVTX_FN();
	if (!c || !J || !codes)
		return fail(c, MC_ERR_ARG, "mc_set_data: null argument");
	CK(cudaSetDevice(c->device));
	int rc = set_dims(c, I, L, P, J);
	if (rc)
		return rc;
	const size_t n = (size_t)I * L * P;
	CK(MC_DEV_MALLOC(&c->d_nat, n));
	CK(cudaMemcpyAsync(c->d_nat, codes, n, cudaMemcpyHostToDevice, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return MC_OK;
}

extern "C" int mc_set_data_synth(mc_ctx *c, int64_t I, int32_t L,
	const mcs_params *g, int64_t i_first)
{
	NVTX_FN();
	if (!c || !g)
		return fail(c, MC_ERR_ARG, "mc_set_data_synth: null argument");
	if (g->jmax < 2 || g->jmax > 254 || g->ploidy < 1 || g->ploidy > 16 || g->K < 1)
		return fail(c, MC_ERR_ARG, "mc_set_data_synth: bad generator parameters");
	CK(cudaSetDevice(c->device));
	free_data(c);
	const int P = g->ploidy;
	const size_t n = (size_t)I * L * P;
	unsigned *d_present = nullptr, *d_missing = nullptr;
	unsigned char *d_map = nullptr;
	CK(MC_DEV_MALLOC(&c->d_nat, n));
	CK(MC_DEV_MALLOC(&d_present, sizeof(unsigned) * 8 * (size_t)L));
	CK(MC_DEV_MALLOC(&d_missing, sizeof(unsigned) * (size_t)L));
	CK(cudaMemsetAsync(d_present, 0, sizeof(unsigned) * 8 * (size_t)L, c->stream));
	CK(cudaMemsetAsync(d_missing, 0, sizeof(unsigned) * (size_t)L, c->stream));
	k_synth_fill<<<grid_for(c, I * (long long)L, 256), 256, 0, c->stream>>>(
		c->d_nat, I, L, *g, i_first, d_present, d_missing);
	LAUNCH_CHECK("k_synth_fill");
	std::vector<unsigned> present((size_t)L * 8), missing((size_t)L);
	CK(cudaMemcpyAsync(present.data(), d_present, sizeof(unsigned) * 8 * (size_t)L,
		cudaMemcpyDeviceToHost, c->stream));
	CK(cudaMemcpyAsync(missing.data(), d_missing, sizeof(unsigned) * (size_t)L,
		cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	/* recode like the reference parser: unobserved alleles get no slot,
	 * a locus with a missing copy gets the phantom slot (read_file.c:527-530) */
	std::vector<int32_t> J((size_t)L);
	std::vector<unsigned char> map((size_t)L * 256, MC_MISSING);
	for (int l = 0; l < L; l++) {
		int nreal = 0;
		for (int code = 0; code < 255; code++)
			if (present[(size_t)l * 8 + (code >> 5)] >> (code & 31) & 1u)
				map[(size_t)l * 256 + code] = (unsigned char)nreal++;
		J[l] = nreal ? nreal + (missing[l] ? 1 : 0) : 0;
	}
	CK(MC_DEV_MALLOC(&d_map, (size_t)L * 256));
	CK(cudaMemcpyAsync(d_map, map.data(), (size_t)L * 256, cudaMemcpyHostToDevice, c->stream));
	k_synth_remap<<<grid_for(c, I * (long long)L, 256), 256, 0, c->stream>>>(
		c->d_nat, I, L, P, d_map);
	LAUNCH_CHECK("k_synth_remap");
	CK(cudaStreamSynchronize(c->stream));
	cudaFree(d_present); cudaFree(d_missing); cudaFree(d_map);
	/* set_dims frees data: keep the generated codes across it */
	unsigned char *keep = c->d_nat;
	c->d_nat = nullptr;
	int rc = set_dims(c, I, L, P, J.data());
	c->d_nat = keep;
	return rc;
}

extern "C" int mc_get_dims(const mc_ctx *c, int64_t *I, int32_t *L, int32_t *P,
	int64_t *T)
{
	if (!c || !c->I)
		return MC_ERR_STATE;
	if (I) *I = c->I;
	if (L) *L = c->L;
	if (P) *P = c->P;
	if (T) *T = c->T;
	return MC_OK;
}

extern "C" int mc_get_J(const mc_ctx *c, int32_t *J)
{
	if (!c || !c->I || !J)
		return MC_ERR_STATE;
	memcpy(J, c->J.data(), sizeof(int32_t) * (size_t)c->L);
	return MC_OK;
}

extern "C" int mc_get_codes(mc_ctx *c, uint8_t *codes)
{
	if (!c || !c->I || !codes)
		return fail(c, MC_ERR_STATE, "mc_get_codes: no data");
	CK(cudaSetDevice(c->device));
	CK(cudaMemcpyAsync(codes, c->d_nat, (size_t)c->I * c->L * c->P,
		cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return MC_OK;
}

/* ------------------------------------------------------------------ plan */

struct HostPlan {
	int W, NG, LW, tile_slots, n_tiles, max_rows;
	std::vector<int> slot_locus, slot_off, slot_J, group_rowbase, group_rows,
		tile_rows;
	size_t smem;
};

static size_t smem_need(int max_rows, int KH, int W, int ks, int IB, int nbuf)
{
	/* + 1: the all-zero row that missing copies read */
	size_t rows = (size_t)(max_rows + 1) * KH * 32 * sizeof(double) * nbuf;
	size_t scr = (size_t)2 * MC_NB * W * ks * IB * KH * sizeof(double);
	return rows + scr + (size_t)W * sizeof(double) + 64;
}

/* Deal loci (sorted by slot count, descending) round-robin to tiles so every
 * tile gets the same mix; inside a tile consecutive slots have similar J, so
 * the LW loci sharing a warp's shared-memory rows waste few of them. */
static void build_tiles(const mc_ctx *c, int W, int NG, int LW, HostPlan &hp)
{
	const int L = c->L;
	hp.W = W; hp.NG = NG; hp.LW = LW;
	hp.tile_slots = W * NG * LW;
	hp.n_tiles = (L + hp.tile_slots - 1) / hp.tile_slots;
	std::vector<int> order((size_t)L);
	for (int l = 0; l < L; l++)
		order[l] = l;
	std::stable_sort(order.begin(), order.end(),
		[&](int a, int b) { return c->J[a] > c->J[b]; });
	const size_t ns = (size_t)hp.n_tiles * hp.tile_slots;
	hp.slot_locus.assign(ns, -1);
	hp.slot_off.assign(ns, 0);
	hp.slot_J.assign(ns, 0);
	for (int r = 0; r < L; r++) {
		const int t = r % hp.n_tiles, s = r / hp.n_tiles, l = order[r];
		hp.slot_locus[(size_t)t * hp.tile_slots + s] = l;
		hp.slot_off[(size_t)t * hp.tile_slots + s] = c->off[l];
		hp.slot_J[(size_t)t * hp.tile_slots + s] = c->J[l];
	}
	hp.group_rowbase.assign((size_t)hp.n_tiles * NG * W, 0);
	hp.group_rows.assign((size_t)hp.n_tiles * NG * W, 1);
	hp.tile_rows.assign((size_t)hp.n_tiles, 0);
	hp.max_rows = 0;
	for (int t = 0; t < hp.n_tiles; t++) {
		int rows = 0;
		for (int gw = 0; gw < NG * W; gw++) {
			int mx = 1;	/* at least one row: missing codes read row 0 */
			for (int lw = 0; lw < LW; lw++)
				mx = std::max(mx, hp.slot_J[(size_t)t * hp.tile_slots + (size_t)gw * LW + lw]);
			hp.group_rowbase[(size_t)t * NG * W + gw] = rows;
			hp.group_rows[(size_t)t * NG * W + gw] = mx;
			rows += mx;
		}
		hp.tile_rows[t] = rows;
		hp.max_rows = std::max(hp.max_rows, rows);
	}
}

template <typename T>
static int upload(mc_ctx *c, T *&dst, const std::vector<T> &src)
{
	CK(MC_DEV_MALLOC(&dst, sizeof(T) * std::max<size_t>(src.size(), 1)));
	CK(cudaMemcpyAsync(dst, src.data(), sizeof(T) * src.size(),
		cudaMemcpyHostToDevice, c->stream));
	return MC_OK;
}

/* Split of a launch into (locus chunks) x (individual chunks) = units for the
 * persistent grid, by an estimate of the launch time in microseconds:
 *   rounds x (tiles of the longest unit x t_tile + t_unit)
 *   + the traffic of the per-chunk partial sums (A_ik: one [I][K] block per
 *     locus chunk, written once and read once; allele sums: one [K][T] block
 *     per individual chunk, zeroed, flushed and summed) at ~5 TB/s.
 * Small problems end up with one tile per CTA (the kernel is latency-bound
 * there), large ones with one unit per CTA slot. */
static void choose_chunks(long long n_ltiles, long long n_itiles, long long sms, double t_tile,
	double apart_us_per_lchunk, double npart_us_per_ichunk, long long max_lchunk_tiles,
	int *best_nl, int *best_ni)
{
	const double t_unit = 4.0;
	double best = 1e300;
	*best_nl = 1;
	*best_ni = 1;
	const long long nl_min = std::max<long long>(1, (n_ltiles + max_lchunk_tiles - 1) / max_lchunk_tiles);
	const long long nl_max = std::min<long long>(n_ltiles, std::max<long long>(nl_min, 4 * sms));
	for (long long nl = nl_min; nl <= nl_max; nl++) {
		const long long lt = (n_ltiles + nl - 1) / nl;
		const long long ni_max = std::min<long long>(n_itiles, std::max<long long>(1, 8 * sms / nl));
		for (long long ni = 1; ni <= ni_max; ni++) {
			const long long it = (n_itiles + ni - 1) / ni;
			const long long rounds = (nl * ni + sms - 1) / sms;
			const double t = (double)rounds * ((double)(lt * it) * t_tile + t_unit)
				+ (double)nl * apart_us_per_lchunk + (double)ni * npart_us_per_ichunk;
			if (t < best * (1.0 - 1e-9)) {
				best = t;
				*best_nl = (int)nl;
				*best_ni = (int)ni;
			}
		}
	}
}


static int raise_smem_limit(mc_ctx *c, const void *fn, size_t smem)
{
	for (auto &e : c->smem_attr)
		if (e.first == fn) {
			if (e.second >= smem)
				return MC_OK;
			CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			e.second = smem;
			return MC_OK;
		}
	CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	c->smem_attr.push_back({ fn, smem });
	return MC_OK;
}

/* ---------------------------------------------- two-pass admixture plan */

static int alloc_outputs(mc_ctx *c, int n_tiles, int n_chunks, int n_units, long long Ipad)
{
	const int K = c->K;
	c->act_tiles = n_tiles; c->act_chunks = n_chunks; c->act_units = n_units;
	c->act_Ipad = Ipad;
	/* room for the chunk counts of a digit-sliced plan on the same buffers */
	n_tiles = std::max(n_tiles, c->dg_min_tiles);
	n_chunks = std::max(n_chunks, c->dg_min_chunks);
	const size_t nN = (size_t)n_chunks * K * std::max<int64_t>(c->T, 1);
	CK(MC_DEV_MALLOC(&c->d_Apart, sizeof(double) * (size_t)n_tiles * Ipad * K));
	CK(MC_DEV_MALLOC(&c->d_Npart, sizeof(double) * nN));
	CK(MC_DEV_MALLOC(&c->d_llpart, sizeof(double) * (size_t)n_units));
	/* + 64: room to cut the buffer into equal slices for up to 64 ranks */
	const size_t xcap = (size_t)K * c->T + 1 + K + 64;
	CK(MC_DEV_MALLOC(&c->d_xbuf, sizeof(double) * xcap));
	CK(MC_DEV_MALLOC(&c->d_red, sizeof(double) * (size_t)RED_BLOCKS * 8));
	CK(cudaMemsetAsync(c->d_Npart, 0, sizeof(double) * nN, c->stream));
	CK(cudaMemsetAsync(c->d_xbuf, 0, sizeof(double) * xcap, c->stream));
	return MC_OK;
}

/* ------------------------------------------------ two-pass plan (admix3) */

/* returns MC_OK with c->use3 set when the kernel of mc_admix3.cuh applies */
static int make_plan3(mc_ctx *c)
{
	c->use3 = false;
	if (c->PP > 8 || c->K > 16 || c->T < 1)
		return MC_OK;
	if (c->opt_kernel == MC_KERNEL_TILE || c->opt_kernel == MC_KERNEL_DENSE)
		return MC_OK;
	const int K = c->K, KP = (K + 1) / 2, KR = 2 * KP, PP = c->PP, L = c->L;
	const int LT = A3_NC / PP;
	const int n_ltiles = (L + LT - 1) / LT;
	const long long n_itiles = (c->I + A3_IT - 1) / A3_IT;
	int cap = c->l3_cap;	/* entries of a tile's list (set when the layout is built) */
	const bool timing = c->opt_timing != 0;
	auto t_prev = std::chrono::steady_clock::now();
	auto mark = [&](const char *what) {
		if (!timing)
			return;
		cudaStreamSynchronize(c->stream);
		const auto t = std::chrono::steady_clock::now();
		fprintf(stderr, "plan3: %-28s %.3f s\n", what,
			std::chrono::duration<double>(t - t_prev).count());
		t_prev = t;
	};
	if (n_itiles * n_ltiles > 0x7fffffffLL)
		return MC_OK;

	/* ---- data-only part: built once per data set, reused for every K ---- */
	if (!c->layout3) {
		/* allele counts: column order of every locus tile */
		unsigned *d_hist = nullptr;
		std::vector<unsigned> hist((size_t)c->T);
		CK(MC_DEV_MALLOC(&d_hist, sizeof(unsigned) * (size_t)c->T));
		CK(cudaMemsetAsync(d_hist, 0, sizeof(unsigned) * (size_t)c->T, c->stream));
		k_allele_hist<<<grid_for(c, c->I * (long long)L, 256), 256, 0, c->stream>>>(
			c->d_nat, c->I, L, c->P, c->d_off, d_hist);
		LAUNCH_CHECK("k_allele_hist");
		CK(cudaMemcpyAsync(hist.data(), d_hist, sizeof(unsigned) * (size_t)c->T,
			cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		cudaFree(d_hist);
		mark("allele histogram");

		/* Rows of a locus inside the kernel.  Pass 1 reads p_s[k][row] with 8-byte
		 * loads, so rows r and r + 16 of a locus share a bank pair and two lanes of a
		 * half warp that carry them cost a wavefront more.  A locus with 17..32
		 * allele slots gets its rows in an order that pairs the rarest alleles with
		 * each other: with m = J - 16, the 16 - m most frequent alleles take rows
		 * m..15 (no partner), the next m rows 0..m-1, and the m rarest rows 16..16+m-1,
		 * the very rarest (the phantom slot of a locus with missing data: no carriers)
		 * opposite row 0, which is also what a missing copy reads. */
		std::vector<unsigned char> perm_of((size_t)c->T);
		std::vector<int> nat_of((size_t)c->T);
		bool permuted = false;
		{
			std::vector<int> rk;
			for (int l = 0; l < L; l++) {
				const int J = c->J[l], o = c->off[l];
				rk.resize((size_t)J);
				for (int j = 0; j < J; j++)
					rk[(size_t)j] = j;
				if (J > 16 && J <= 32) {
					permuted = true;
					std::stable_sort(rk.begin(), rk.end(), [&](int x, int y) {
						return hist[(size_t)o + x] > hist[(size_t)o + y]; });
					const int m = J - 16;
					std::vector<int> at((size_t)J);	/* row -> allele */
					for (int r = m; r < 16; r++)
						at[(size_t)r] = rk[(size_t)(r - m)];
					for (int k = 0; k < m; k++) {
						at[(size_t)k] = rk[(size_t)(16 - m + k)];
						at[(size_t)(16 + k)] = rk[(size_t)(J - 1 - k)];
					}
					rk = at;
				}
				for (int r = 0; r < J; r++) {
					perm_of[(size_t)o + rk[(size_t)r]] = (unsigned char)r;
					nat_of[(size_t)o + r] = o + rk[(size_t)r];
				}
			}
		}
		std::vector<int> lt_ncol((size_t)n_ltiles);
		std::vector<std::vector<std::pair<unsigned, unsigned short>>> cols((size_t)n_ltiles);
		int ncm = 1, mtr = 1;
		for (int lt = 0; lt < n_ltiles; lt++) {
			const int lf = lt * LT, le = std::min(L, lf + LT);
			mtr = std::max(mtr, c->off[le] - c->off[lf]);
			auto &v = cols[lt];
			for (int l = lf; l < le; l++)
				for (int j = 0; j < c->J[l]; j++)
					if (hist[(size_t)c->off[l] + j])
						v.push_back({ hist[(size_t)c->off[l] + j],
							(unsigned short)((l - lf) << 8
								| perm_of[(size_t)c->off[l] + j]) });
			std::stable_sort(v.begin(), v.end(),
				[](const std::pair<unsigned, unsigned short> &x,
				   const std::pair<unsigned, unsigned short> &y) { return x.first > y.first; });
			lt_ncol[lt] = (int)v.size();
			ncm = std::max(ncm, (int)v.size());
		}
		if (ncm > A3_THREADS || ncm >= 255) {
			free_layout3(c);
			return MC_OK;	/* more allele columns in a tile than lanes */
		}
		/* most frequent allele first; the lanes per column are chosen per
		 * tile by k3_build_csc */
		std::vector<unsigned short> colinfo((size_t)n_ltiles * ncm, 0);
		for (int lt = 0; lt < n_ltiles; lt++)
			for (size_t x = 0; x < cols[lt].size(); x++)
				colinfo[(size_t)lt * ncm + x] = cols[lt][x].second;
		free_layout3(c);
		int rc;
		if ((rc = upload(c, c->d3_lt_ncol, lt_ncol))) return rc;
		if ((rc = upload(c, c->d3_colinfo, colinfo))) return rc;
		if ((rc = upload(c, c->d3_perm_of, perm_of))) return rc;
		if ((rc = upload(c, c->d3_nat_of, nat_of))) return rc;
		const size_t ntile = (size_t)n_itiles * n_ltiles;
		mark("column order + uploads");
		CK(MC_DEV_MALLOC(&c->d3_codes, ntile * A3_THREADS * A3_NC));
		CK(MC_DEV_MALLOC(&c->d3_colstart, ntile * (3 * (size_t)((ncm + 1 + 7) / 8 * 8)
			+ A3_THREADS) * sizeof(unsigned short)));
		mark("cudaMalloc codes/tables");
		k3_build_codes<<<grid_for(c, (long long)ntile * A3_THREADS, 256), 256, 0, c->stream>>>(
			c->d_nat, c->d3_codes, c->I, L, c->P, PP, (int)n_itiles, n_ltiles,
			c->d_off, c->d3_perm_of);
		LAUNCH_CHECK("k3_build_codes");
		mark("k3_build_codes");
		/* the lists are stored step-major, [pair of steps][thread][2]: their length is
		 * the longest lane list of any tile (an even number of steps) x A3_THREADS, known
		 * after a first, counting-only launch of the builder */
		const size_t bsm = a3_build_smem_bytes(ncm);
		int qmax = 0;
		CK(cudaMemsetAsync(c->d_small, 0, sizeof(int), c->stream));
		k3_build_csc<<<(unsigned)ntile, 128, bsm, c->stream>>>(c->d3_codes, PP, n_ltiles,
			ncm, 0, c->d3_lt_ncol, c->d3_colinfo, nullptr, c->d3_colstart,
			reinterpret_cast<int *>(c->d_small));
		LAUNCH_CHECK("k3_build_csc (count)");
		CK(cudaMemcpyAsync(&qmax, c->d_small, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		cap = (std::max(qmax, 1) + 1) / 2 * 2 * A3_THREADS;
		mark("k3_build_csc (count)");
		CK(MC_DEV_MALLOC(&c->d3_csc, ntile * (size_t)cap * sizeof(unsigned short)));
		k3_build_csc<<<(unsigned)ntile, 128, bsm, c->stream>>>(c->d3_codes, PP, n_ltiles,
			ncm, cap, c->d3_lt_ncol, c->d3_colinfo, c->d3_csc, c->d3_colstart, nullptr);
		LAUNCH_CHECK("k3_build_csc");
		mark("k3_build_csc");
		c->l3_cap = cap;
		c->l3_permuted = permuted;
		c->l3_ncolmax = ncm;
		c->l3_max_tile_rows = mtr;
		c->layout3 = true;
	}
	const int ncolmax = c->l3_ncolmax, max_tile_rows = c->l3_max_tile_rows;

	/* shared memory holds the tile only; the chunk's allele sums live in L2 */
	c->smem3 = a3_smem_bytes(KP, c->admixture ? A3_ADMIX_EM : A3_MIX_M, ncolmax, cap);
	c->smem3_ll = a3_smem_bytes(KP, c->admixture ? A3_ADMIX_LL : A3_MIX_E, ncolmax, cap);
	/* two CTAs per SM while the tile fits twice (K <= 10 at 16 copies per tile),
	 * else one */
	const size_t smem_cap2 = (size_t)(228 * 1024) / A3_CTAS_PER_SM - 1024 - 64;
	const size_t smem_cap1 = (size_t)(227 * 1024) - 64;
	if (c->smem3 + 64 > smem_cap1 || max_tile_rows > A3_PR) {
		free_layout3(c);	/* the one-pass kernel takes over: drop the multi-GB lists */
		return MC_OK;
	}
	const long long sms = (long long)c->num_sms * (c->smem3 + 64 > smem_cap2 ? 1 : A3_CTAS_PER_SM);

	/* locus chunks x individual chunks: ~13 us per (256 individuals x 16 copies)
	 * tile with two CTAs per SM */
	int n_lchunks = 1, n_ichunks = 1;
	choose_chunks(n_ltiles, n_itiles, sms, 13.0,
		(double)n_itiles * A3_IT * K * 16.0 / 5e6,
		(double)K * (double)c->T * 40.0 / 5e6, n_ltiles, &n_lchunks, &n_ichunks);
	std::vector<int> lc_first((size_t)n_lchunks + 1);
	for (int x = 0; x <= n_lchunks; x++)
		lc_first[x] = (int)((long long)n_ltiles * x / n_lchunks);

	Admix3Args &a = c->a3;
	memset(&a, 0, sizeof a);
	a.K = K;
	a.n_itiles = (int)n_itiles; a.n_ltiles = n_ltiles; a.n_lchunks = n_lchunks;
	a.n_ichunks = n_ichunks; a.n_units = n_lchunks * n_ichunks;
	a.I = c->I; a.Ipad = n_itiles * A3_IT; a.T = c->T; a.L = L;
	a.ncolmax = ncolmax; a.max_chunk_rows = 0; a.cap = cap;
	c->KP3 = KP;
	c->grid3 = (int)std::min<long long>(a.n_units, sms);

	int rc;
	if ((rc = upload(c, c->d3_lc_first, lc_first))) return rc;
	a.lt_ncol = c->d3_lt_ncol; a.colinfo = c->d3_colinfo;
	 a.lc_first = c->d3_lc_first; a.off = c->d_off;
	a.nat_of = c->l3_permuted ? c->d3_nat_of : nullptr;
	a.codes = c->d3_codes; a.csc = c->d3_csc; a.colstart = c->d3_colstart;
	if ((rc = alloc_outputs(c, n_lchunks, n_ichunks, a.n_units, a.Ipad))) return rc;
	CK(MC_DEV_MALLOC(&c->d3_Gacc, sizeof(double) * (size_t)n_ichunks * c->T * KR));
	if (c->l3_permuted)
		CK(MC_DEV_MALLOC(&c->d3_pperm, sizeof(double) * (size_t)c->K * c->T));
	a.Apart = c->d_Apart; a.Npart = c->d_Npart; a.llpart = c->d_llpart;
	a.Gacc = c->d3_Gacc;
	CK(cudaStreamSynchronize(c->stream));
	mark("partial-sum buffers");
	c->use3 = true;
	return MC_OK;
}

/* mode: A3_ADMIX_EM / A3_ADMIX_LL / A3_MIX_E (p = the log p table) / A3_MIX_M (eta = v_ik) */
static int launch_admix3(mc_ctx *c, int mode, const double *p, const double *eta,
	long long eta_stride, bool fallback = false)
{
	admix3_fn fn = mc_pick_admix3(mode, c->KP3, c->PP);
	if (!fn)
		return fail(c, MC_ERR_UNSUPPORTED, "no admix3 kernel for K=%d P=%d", c->K, c->P);
	Admix3Args a = c->a3;
	a.p = a.p_nat = p; a.eta = eta; a.eta_stride = eta_stride;
	if (a.nat_of && p) {	/* the rows of some loci are reordered inside the kernel */
		k3_permute_rows<<<grid_for(c, (long long)c->K * c->T, 256), 256, 0, c->stream>>>(
			p, c->d3_pperm, a.nat_of, c->K, c->T);
		LAUNCH_CHECK("k3_permute_rows");
		a.p = c->d3_pperm;
	}
	if (fallback) {	/* behind the digit kernels: runs only when they declined */
		a.run_if = c->d_dg_flag;
		a.n_chunks_dev = c->d_dg_flag + 1;
	}
	const size_t smem = a3_smem_bytes(c->KP3, mode, a.ncolmax, a.cap);
	{
		const int rca = raise_smem_limit(c, (const void *)fn, smem);
		if (rca)
			return rca;
	}
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	if (c->profile) {
		CK(cudaEventCreate(&e0));
		CK(cudaEventCreate(&e1));
		CK(cudaEventRecord(e0, c->stream));
	}
	fn<<<c->grid3, A3_THREADS, smem, c->stream>>>(a);
	LAUNCH_CHECK(mode == A3_ADMIX_EM ? "admix3_kernel<ADMIX_EM>" : mode == A3_ADMIX_LL
		? "admix3_kernel<ADMIX_LL>" : mode == A3_MIX_E ? "admix3_kernel<MIX_E>"
		: "admix3_kernel<MIX_M>");
	if (c->profile) {
		CK(cudaEventRecord(e1, c->stream));
		c->prof_events.push_back({ e0, e1 });
	}
	return MC_OK;
}


/* ------------------------------------------------ dense DMMA plan (mc_dense) */

/* ---------------------------------------------- digit-sliced mixture plan */

/* Chunks of the contraction dimension for digit_kernel: n_ctarows CTAs per
 * chunk, `slots` CTAs resident at once, ~t_block_us per CTA and block, and
 * partial_us for writing and re-reading one chunk's partial sums. */
static int digit_chunks(long long n_ctarows, int n_blocks, long long slots, int min_chunks,
	double t_block_us, double partial_us)
{
	int best = std::max(1, std::min(min_chunks, n_blocks));
	double best_cost = 1e300;
	const int hi = std::max(best, std::min(n_blocks, 256));
	for (int nc = best; nc <= hi; nc++) {
		const double waves = (double)((n_ctarows * nc + slots - 1) / slots);
		const double blocks = (double)((n_blocks + nc - 1) / nc);
		const double cost = waves * (blocks + 2.0) * t_block_us + nc * partial_us;
		if (cost < best_cost - 1e-9) {
			best_cost = cost;
			best = nc;
		}
	}
	return best;
}

/* called from make_plan_dense for the mixture model: the same biallelic data,
 * K <= 16.  Sets c->use_dg and the chunk counts of the two passes. */
static int make_plan_digit(mc_ctx *c, int general, long long Ipad, int *ncE, int *ncM)
{
	c->use_dg = false;
	*ncE = *ncM = 0;
	if (c->admixture || c->opt_kernel == MC_KERNEL_DENSE || c->K > 16 || c->P > 15 || c->T < 1)
		return MC_OK;
	const int K = c->K, R = dg_R(K);
	digit_fn fE = mc_pick_digit(K, DG_MIX_E), fM = mc_pick_digit(K, DG_MIX_M);
	if (!fE || !fM)
		return MC_OK;
	/* bytes per row of the count matrix: one per locus, or one per column pair */
	const long long nb = general ? (c->T + 1) / 2 : c->L;
	const long long mtE = (c->I + 15) / 16, blE = (nb + 63) / 64;
	const long long mtM = (nb + 7) / 8, blM = (c->I + 127) / 128;
	if (mtE * blE > 0x7fffffffLL || mtM * blM > 0x7fffffffLL || blM > 0x7fffffffLL)
		return MC_OK;
	if (c->layout_dg && c->dg_kind != (general ? 2 : 1))
		return MC_OK;
	if (!c->layout_dg) {
		const size_t need = sizeof(uint4) * ((size_t)mtE * blE + (size_t)mtM * blM) * 64;
		size_t mfree = 0, mtotal = 0;
		CK(cudaMemGetInfo(&mfree, &mtotal));
		/* the two count layouts must leave room for the model: else the
		 * gather kernels run the mixture passes */
		if (general && need + (need >> 2) + ((size_t)1 << 30) > mfree)
			return MC_OK;
		if (general) {
			std::vector<int> cl((size_t)c->T);
			for (int l = 0; l < c->L; l++)
				for (int j = 0; j < c->J[l]; j++)
					cl[(size_t)c->off[l] + j] = l;
			int rcu;
			if ((rcu = upload(c, c->d_col_locus, cl))) return rcu;
		}
		CK(MC_DEV_MALLOC(&c->d_dg_cntE, sizeof(uint4) * (size_t)mtE * blE * 64));
		CK(MC_DEV_MALLOC(&c->d_dg_cntM, sizeof(uint4) * (size_t)mtM * blM * 64));
		k_digit_counts<<<grid_for(c, mtE * blE * 64, 256), 256, 0, c->stream>>>(c->d_nat,
			c->d_dg_cntE, c->I, c->L, c->P, (int)mtE, (int)blE, DG_MIX_E, general, c->T,
			c->d_col_locus, c->d_off);
		LAUNCH_CHECK("k_digit_counts");
		k_digit_counts<<<grid_for(c, mtM * blM * 64, 256), 256, 0, c->stream>>>(c->d_nat,
			c->d_dg_cntM, c->I, c->L, c->P, (int)mtM, (int)blM, DG_MIX_M, general, c->T,
			c->d_col_locus, c->d_off);
		LAUNCH_CHECK("k_digit_counts");
		c->layout_dg = true;
		c->dg_kind = general ? 2 : 1;
	}
	CK(MC_DEV_MALLOC(&c->d_dg_tabE, sizeof(uint2) * (size_t)blE * 4 * K * 32));
	CK(MC_DEV_MALLOC(&c->d_dg_tabM, sizeof(uint2) * (size_t)blM * 4 * K * 32));
	CK(MC_DEV_MALLOC(&c->d_dg_flag, sizeof(int) * 2));
	CK(cudaMemsetAsync(c->d_dg_flag, 0, sizeof(int) * 2, c->stream));
	CK(MC_DEV_MALLOC(&c->d_dg_vscale, sizeof(double) * 2 * K));
	CK(cudaMemsetAsync(c->d_dg_vscale, 0, sizeof(double) * 2 * K, c->stream));

	/* 32-bit accumulators: count x digit <= 255 P per element of the
	 * contraction dimension */
	const long long max_len = 0x7fffffffLL / (255LL * std::max(c->P, 1));
	int rcs;
	if ((rcs = raise_smem_limit(c, (const void *)fE, dg_smem_bytes(K)))) return rcs;
	if ((rcs = raise_smem_limit(c, (const void *)fM, dg_smem_bytes(K)))) return rcs;
	auto plan = [&](DigitArgs &a, digit_fn fn, long long mt, long long bl, int per_block,
		double partial_us) {
		int occ = 1;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)fn, DG_THREADS,
			dg_smem_bytes(K));
		occ = std::max(occ, 1);
		memset(&a, 0, sizeof a);
		a.K = K; a.L = c->L; a.general = general;
		a.n_mtiles = (int)mt; a.n_blocks = (int)bl;
		const long long wunits = (mt + R - 1) / R;
		a.n_ctarows = (int)((wunits + DG_WARPS - 1) / DG_WARPS);
		const int min_chunks = (int)((bl * per_block + max_len - 1) / max_len);
		/* a CTA and block: 4 warps x 4 steps x R x K IMMAs at 2 clk each on
		 * an SM shared by `occ` CTAs, or 16 R KB of counts at the SM's share
		 * of HBM, whichever is longer */
		const double t_mma = DG_WARPS * 4.0 * R * K * 2.0 * occ / 1900.0;
		const double t_hbm = DG_WARPS * R * 1024.0 * occ / 43000.0;
		a.n_chunks = digit_chunks(a.n_ctarows, (int)bl, (long long)c->num_sms * occ,
			min_chunks, std::max(t_mma, t_hbm), partial_us);
		a.I = c->I; a.Ipad = Ipad; a.T = c->T;
		a.off = c->d_off; a.J = c->d_J;
	};
	/* elements of the contraction dimension per block that can carry counts:
	 * 64 loci, or 128 columns that may be 128 loci, or 128 individuals */
	plan(c->dgE, fE, mtE, blE, general ? 128 : 64, (double)c->I * K * 16.0 / 5e6);
	plan(c->dgM, fM, mtM, blM, 128, (double)K * (double)c->T * 16.0 / 5e6);
	c->dgE.cnt = c->d_dg_cntE; c->dgE.tab = c->d_dg_tabE;
	c->dgM.cnt = c->d_dg_cntM; c->dgM.tab = c->d_dg_tabM;
	if ((long long)c->dgE.n_chunks * c->dgE.n_ctarows > 0x7fffffffLL
		|| (long long)c->dgM.n_chunks * c->dgM.n_ctarows > 0x7fffffffLL)
		return MC_OK;
	*ncE = c->dgE.n_chunks;
	*ncM = c->dgM.n_chunks;
	c->dg_min_tiles = *ncE;
	c->dg_min_chunks = *ncM;
	c->use_dg = true;
	return MC_OK;
}

/* One pass of the digit-sliced kernels.  DG_MIX_E: src = p [K][T], take_log as
 * in k_digit_table (2 = the log-likelihood pass, where log 0 makes the pass
 * fall back: the caller launches the FP64 kernels behind it, gated on the
 * flag); DG_MIX_M: src = posteriors [I][K]. */
static int launch_digit(mc_ctx *c, int mode, const double *src, int take_log)
{
	digit_fn fn = mc_pick_digit(c->K, mode);
	if (!fn)
		return fail(c, MC_ERR_UNSUPPORTED, "no digit kernel for K=%d", c->K);
	DigitArgs a = mode == DG_MIX_E ? c->dgE : c->dgM;
	const long long n = (long long)a.n_blocks * 4 * c->K;
	k_digit_table<<<grid_for(c, n * 32, 256), 256, 0, c->stream>>>(src, c->d_dg_vscale,
		const_cast<uint2 *>(a.tab), c->d_dg_flag, mode == DG_MIX_E ? take_log : 0, c->K,
		c->I, c->L, c->T, c->d_off, c->d_J, a.n_blocks, a.general);
	LAUNCH_CHECK("k_digit_table");
	a.out = mode == DG_MIX_E ? c->d_Apart : c->d_Npart;
	a.Ipad = c->act_Ipad;
	a.unscale = c->d_dg_vscale + c->K;
	if (mode == DG_MIX_E && take_log == 2) {
		a.skip_if = c->d_dg_flag;
		a.n_chunks_dev = c->d_dg_flag + 1;
	}
	{
		const int rca = raise_smem_limit(c, (const void *)fn, dg_smem_bytes(c->K));
		if (rca)
			return rca;
	}
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	if (c->profile) {
		CK(cudaEventCreate(&e0));
		CK(cudaEventCreate(&e1));
		CK(cudaEventRecord(e0, c->stream));
	}
	fn<<<(unsigned)(a.n_chunks * a.n_ctarows), DG_THREADS, dg_smem_bytes(c->K), c->stream>>>(a);
	LAUNCH_CHECK("digit_kernel");
	if (c->profile) {
		CK(cudaEventRecord(e1, c->stream));
		c->prof_events.push_back({ e0, e1 });
	}
	return MC_OK;
}

/* returns MC_OK with c->use_dn set when the kernels of mc_dense.cuh apply:
 * no allele code above 1 anywhere (every locus has at most two observed
 * alleles; a third slot can only be the phantom slot of read_file.c:527-530,
 * which no copy carries), K <= 16, ploidy <= 15 */
static int make_plan_dense(mc_ctx *c)
{
	c->use_dn = false;
	if (c->opt_kernel != MC_KERNEL_AUTO && c->opt_kernel != MC_KERNEL_DENSE
		&& c->opt_kernel != MC_KERNEL_DIGIT)
		return MC_OK;
	if (c->K > 16 || c->P > 15 || c->T < 1)
		return MC_OK;
	for (int l = 0; l < c->L; l++)
		if (c->J[l] > 3)
			return MC_OK;
	if (c->dn_maxcode < 0) {
		unsigned *d_m = nullptr, m = 0;
		CK(MC_DEV_MALLOC(&d_m, sizeof(unsigned)));
		CK(cudaMemsetAsync(d_m, 0, sizeof(unsigned), c->stream));
		const long long n = c->I * (long long)c->L * c->P;
		k_dense_maxcode<<<grid_for(c, n, 256), 256, 0, c->stream>>>(c->d_nat, n, d_m);
		LAUNCH_CHECK("k_dense_maxcode");
		CK(cudaMemcpyAsync(&m, d_m, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		cudaFree(d_m);
		c->dn_maxcode = (int)m;
	}
	if (c->dn_maxcode > 1)
		return MC_OK;

	const int NB = c->K <= 8 ? 1 : 2;
	const long long n_itiles = (c->I + DN_IT - 1) / DN_IT;
	const int n_ltiles = (c->L + DN_TL - 1) / DN_TL;
	if (n_itiles * n_ltiles > 0x7fffffffLL)
		return MC_OK;
	if (!c->layout_dn) {
		CK(MC_DEV_MALLOC(&c->d_dn_cnt, (size_t)n_itiles * n_ltiles * DN_IT * 16));
		k_dense_counts<<<grid_for(c, n_itiles * n_ltiles * DN_IT, 256), 256, 0, c->stream>>>(
			c->d_nat, c->d_dn_cnt, c->I, c->L, c->P, (int)n_itiles, n_ltiles);
		LAUNCH_CHECK("k_dense_counts");
		c->layout_dn = true;
	}
	/* shared memory: fixed part, the rest holds the chunk's allele sums */
	const size_t fixed = dn_smem_bytes(NB, DN_ADMIX_EM, 0) + 64;
	const size_t smem_cap = (size_t)(228 * 1024) / DN_CTAS_PER_SM - 1024 - 64;
	const long long budget = (long long)((smem_cap - fixed) / ((size_t)NB * 256 * sizeof(double)));
	if (budget < 1)
		return MC_OK;
	const long long sms = (long long)c->num_sms * DN_CTAS_PER_SM;
	int best_nl = 1, best_ni = 1;
	/* ~7 us per (256 individuals x 16 loci) tile with two CTAs per SM */
	choose_chunks(n_ltiles, n_itiles, sms, 7.0,
		(double)n_itiles * DN_IT * c->K * 16.0 / 5e6,
		(double)c->K * (double)c->T * 24.0 / 5e6, budget, &best_nl, &best_ni);
	std::vector<int> lc_first((size_t)best_nl + 1);
	int max_chunk_tiles = 1;
	for (int x = 0; x <= best_nl; x++)
		lc_first[x] = (int)((long long)n_ltiles * x / best_nl);
	for (int x = 0; x < best_nl; x++)
		max_chunk_tiles = std::max(max_chunk_tiles, lc_first[x + 1] - lc_first[x]);

	DenseArgs &a = c->dn;
	memset(&a, 0, sizeof a);
	a.K = c->K; a.L = c->L;
	a.n_itiles = (int)n_itiles; a.n_ltiles = n_ltiles;
	a.n_lchunks = best_nl; a.n_ichunks = best_ni; a.n_units = best_nl * best_ni;
	a.max_chunk_tiles = max_chunk_tiles;
	a.I = c->I; a.Ipad = n_itiles * DN_IT; a.T = c->T;
	c->dn_NB = NB;
	/* largest allele count the log-likelihood powers must handle */
	c->dn_pbits = c->P <= 1 ? 1 : c->P <= 3 ? 2 : c->P == 4 ? 4 : c->P <= 7 ? 7 : 15;
	c->grid_dn = (int)std::min<long long>(a.n_units, sms);
	int rc;
	if ((rc = upload(c, c->d_dn_lc_first, lc_first))) return rc;
	CK(MC_DEV_MALLOC(&c->d_dn_pd, sizeof(double) * (size_t)n_ltiles * DN_TL * dn_pl(NB)));
	a.lc_first = c->d_dn_lc_first; a.off = c->d_off; a.J = c->d_J;
	a.cnt = c->d_dn_cnt; a.pd = c->d_dn_pd;
	int ncE = 0, ncM = 0;
	if ((rc = make_plan_digit(c, 0, a.Ipad, &ncE, &ncM))) return rc;
	if ((rc = alloc_outputs(c, std::max(best_nl, ncE), std::max(best_ni, ncM), a.n_units,
		a.Ipad))) return rc;
	c->dn_nl = best_nl; c->dn_ni = best_ni;
	c->act_tiles = c->use_dg ? ncE : best_nl;
	c->act_chunks = c->use_dg ? ncM : best_ni;
	a.Apart = c->d_Apart; a.Npart = c->d_Npart; a.llpart = c->d_llpart;
	CK(cudaStreamSynchronize(c->stream));
	c->use_dn = true;
	return MC_OK;
}

/* `ptab`: the [K][T] table the dense p fragments are built from (p; its log,
 * take_log as in k_dense_p, for the mixture E pass; nullptr for the M pass) */
static int launch_dense(mc_ctx *c, int mode, const double *ptab, const double *p,
	const double *eta, long long eta_stride, int take_log = 0, bool fallback = false)
{
	dense_fn fn = mc_pick_dense(c->dn_NB, c->dn_pbits, mode);
	if (!fn)
		return fail(c, MC_ERR_UNSUPPORTED, "no dense kernel for K=%d P=%d", c->K, c->P);
	DenseArgs a = c->dn;
	a.p = p; a.eta = eta; a.eta_stride = eta_stride;
	if (fallback) {	/* behind the digit kernels: runs only when they declined */
		a.run_if = c->d_dg_flag;
		a.n_chunks_dev = c->d_dg_flag + 1;
	}
	if (ptab) {
		const int K8 = 8 * c->dn_NB, npad = a.n_ltiles * DN_TL;
		k_dense_p<<<grid_for(c, (long long)npad * K8 * 2, 256), 256, 0, c->stream>>>(ptab,
			c->d_dn_pd, c->d_off, c->d_J, c->K, c->L, c->T, npad, dn_pl(c->dn_NB), K8,
			take_log, a.run_if);
		LAUNCH_CHECK("k_dense_p");
	}
	const size_t smem = dn_smem_bytes(c->dn_NB, mode, a.max_chunk_tiles);
	{
		const int rca = raise_smem_limit(c, (const void *)fn, smem);
		if (rca)
			return rca;
	}
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	if (c->profile) {
		CK(cudaEventCreate(&e0));
		CK(cudaEventCreate(&e1));
		CK(cudaEventRecord(e0, c->stream));
	}
	fn<<<c->grid_dn, DN_THREADS, smem, c->stream>>>(a);
	LAUNCH_CHECK("dense_kernel");
	if (c->profile) {
		CK(cudaEventRecord(e1, c->stream));
		c->prof_events.push_back({ e0, e1 });
	}
	return MC_OK;
}

static int make_plan_base(mc_ctx *c);

/* The planner: dense / digit-sliced plans for biallelic data; else, for the
 * mixture model, the digit-sliced plan on column pairs (mc_digit.cuh) on top
 * of the gather plan that stays as its fall-back; else the gather plans. */
static int make_plan(mc_ctx *c)
{
	{
		const int rcd = make_plan_dense(c);
		if (rcd || c->use_dn)
			return rcd;
	}
	int ncE = 0, ncM = 0;
	if (!c->admixture && (c->opt_kernel == MC_KERNEL_AUTO || c->opt_kernel == MC_KERNEL_DIGIT)
		&& c->dg_kind != 1) {
		const int rcg = make_plan_digit(c, 1, 0, &ncE, &ncM);
		if (rcg)
			return rcg;
	}
	const bool dg = c->use_dg;
	const int rc = make_plan_base(c);	/* free of use_dg; sizes the partial sums */
	if (rc)
		return rc;
	if (dg) {
		c->use_dg = true;
		c->act_tiles = ncE;
		c->act_chunks = ncM;
	}
	return MC_OK;
}

static int make_plan_base(mc_ctx *c)
{
	const int K = c->K;
	{
		const int rc3 = make_plan3(c);
		if (rc3 || c->use3)
			return rc3;
	}
	int ks = 1;
	const int kh_max = KH_MAX;
	while ((K + ks - 1) / ks > kh_max && ks < 32)
		ks <<= 1;
	const int KH = (K + ks - 1) / ks;
	if (KH > KH_MAX)
		return fail(c, MC_ERR_UNSUPPORTED, "K=%d exceeds the supported "
			"maximum of %d clusters", K, KH_MAX * 32);
	const int LW = 32 / ks;
	const int nbuf = 2;	/* p rows + accumulator rows */
	c->k_split = ks;
	c->KH = KH;

	/* largest tile that fits shared memory; prefer many warps */
	HostPlan best;
	bool found = false;
	const int max_slots = ((c->L + LW - 1) / LW) * LW;
	for (int W = 8; W >= 1 && !found; W--) {
		int NG = std::max(1, std::min(64, max_slots / (W * LW)));
		for (; NG >= 1; NG--) {
			/* do not build tiles much larger than the data needs */
			if (NG > 1 && (long long)W * (NG - 1) * LW >= c->L)
				continue;
			HostPlan hp;
			build_tiles(c, W, NG, LW, hp);
			hp.smem = smem_need(hp.max_rows, KH, W, ks, c->IB, nbuf);
			if (hp.smem <= SMEM_LIMIT) {
				best = std::move(hp);
				found = true;
				break;
			}
		}
	}
	if (!found)
		return fail(c, MC_ERR_UNSUPPORTED, "a single locus group needs more "
			"than %d bytes of shared memory (K=%d)", SMEM_LIMIT, K);

	const long long n_blocks = (c->I + c->IB - 1) / c->IB;
	/* chunks of individuals: enough units to balance the persistent grid */
	int C = 1;
	{
		const long long sms = c->num_sms;
		double best_eff = -1;
		const long long cmax = std::min<long long>(n_blocks,
			std::max<long long>(1, (8 * sms + best.n_tiles - 1) / best.n_tiles));
		for (long long cc = 1; cc <= cmax; cc++) {
			const long long units = cc * best.n_tiles;
			const long long rounds = (units + sms - 1) / sms;
			double eff = (double)units / (double)(rounds * sms);
			/* mild preference for fewer chunks (fewer partial sums) */
			eff -= 0.002 * (double)cc;
			if (eff > best_eff + 1e-12) {
				best_eff = eff;
				C = (int)cc;
			}
		}
	}

	TileArgs &ta = c->ta;
	memset(&ta, 0, sizeof ta);
	ta.K = K; ta.k_split = ks; ta.loci_per_warp = LW; ta.warps = best.W;
	ta.groups = best.NG; ta.n_tiles = best.n_tiles; ta.n_chunks = C;
	ta.n_units = best.n_tiles * C; ta.tile_slots = best.tile_slots;
	ta.n_blocks = n_blocks; ta.I = c->I; ta.Ipad = n_blocks * c->IB; ta.T = c->T;
	ta.max_rows = best.max_rows;
	const int UB = c->IB * c->PP;
	ta.tile_stride = (long long)n_blocks * best.tile_slots * UB;
	c->smem_em = best.smem;
	c->block = best.W * 32;
	c->grid = std::min(ta.n_units, c->num_sms);

	int rc;
	if ((rc = upload(c, c->d_slot_locus, best.slot_locus))) return rc;
	if ((rc = upload(c, c->d_slot_off, best.slot_off))) return rc;
	if ((rc = upload(c, c->d_slot_J, best.slot_J))) return rc;
	if ((rc = upload(c, c->d_group_rowbase, best.group_rowbase))) return rc;
	if ((rc = upload(c, c->d_group_rows, best.group_rows))) return rc;
	if ((rc = upload(c, c->d_tile_rows, best.tile_rows))) return rc;
	ta.slot_locus = c->d_slot_locus; ta.slot_off = c->d_slot_off;
	ta.slot_J = c->d_slot_J; ta.group_rowbase = c->d_group_rowbase;
	ta.group_rows = c->d_group_rows; ta.tile_rows = c->d_tile_rows;

	CK(MC_DEV_MALLOC(&c->d_tiled, (size_t)ta.tile_stride * best.n_tiles));
	k_tile_codes<<<grid_for(c, (long long)best.n_tiles * n_blocks * best.tile_slots, 256),
		256, 0, c->stream>>>(c->d_nat, c->d_tiled, c->d_slot_locus, best.n_tiles,
		best.tile_slots, n_blocks, c->I, c->L, c->P, c->PP, c->IB, ta.tile_stride);
	LAUNCH_CHECK("k_tile_codes");
	ta.codes = c->d_tiled;

	if ((rc = alloc_outputs(c, best.n_tiles, C, ta.n_units, ta.Ipad))) return rc;
	ta.Apart = c->d_Apart; ta.Npart = c->d_Npart; ta.llpart = c->d_llpart;
	CK(cudaStreamSynchronize(c->stream));
	return MC_OK;
}

/* ----------------------------------------------------------------- model */

extern "C" int mc_alloc_model(mc_ctx *c, int32_t K, int admixture,
	int eta_constrained, int q, double eta_lb, double p_lb, int do_projection)
{
	NVTX_FN();
	if (!c)
		return MC_ERR_ARG;
	if (!c->I)
		return fail(c, MC_ERR_STATE, "mc_alloc_model: no data loaded");
	if (K < 1 || q < 0 || q > MC_QMAX)
		return fail(c, MC_ERR_ARG, "mc_alloc_model: bad K=%d or q=%d", K, q);
	CK(cudaSetDevice(c->device));
	free_model(c);
	c->K = K; c->admixture = admixture ? 1 : 0;
	c->eta_constrained = eta_constrained ? 1 : 0;
	c->per_indiv = admixture && !eta_constrained;
	c->q = q; c->eta_lb = eta_lb; c->p_lb = p_lb; c->do_proj = do_projection ? 1 : 0;
	c->neta = c->per_indiv ? c->I * K : K;
	c->np = (int64_t)K * c->T;
	const size_t npb = sizeof(double) * (size_t)std::max<int64_t>(c->np, 1);
	const size_t neb = sizeof(double) * (size_t)c->neta;
	for (int s = 0; s < 3; s++) {
		CK(MC_DEV_MALLOC(&c->d_p[s], npb));
		CK(MC_DEV_MALLOC(&c->d_eta[s], neb));
		CK(cudaMemsetAsync(c->d_p[s], 0, npb, c->stream));
		CK(cudaMemsetAsync(c->d_eta[s], 0, neb, c->stream));
	}
	for (int s = 0; s < q; s++) {
		double *a, *b, *d, *e;
		CK(MC_DEV_MALLOC(&a, npb)); CK(MC_DEV_MALLOC(&b, npb));
		CK(MC_DEV_MALLOC(&d, neb)); CK(MC_DEV_MALLOC(&e, neb));
		c->d_up.push_back(a); c->d_vp.push_back(b);
		c->d_ue.push_back(d); c->d_ve.push_back(e);
	}
	CK(MC_DEV_MALLOC(&c->d_post, sizeof(double) * (size_t)c->I * K));
	CK(cudaMemsetAsync(c->d_post, 0, sizeof(double) * (size_t)c->I * K, c->stream));
	CK(MC_DEV_MALLOC(&c->d_IK, sizeof(int) * (size_t)c->I));
	if (!c->admixture) {
		c->lli_n = (long long)c->num_sms * 4 * (2 * K + 1);
		CK(MC_DEV_MALLOC(&c->d_lli, sizeof(double) * (size_t)c->lli_n));
		CK(MC_DEV_MALLOC(&c->d_logp, npb));
	}
	int rc = make_plan(c);
	if (rc)
		return rc;
	c->have_model = true;
	return MC_OK;
}

extern "C" int mc_eta_len(const mc_ctx *c, int64_t *n)
{
	if (!c || !c->have_model)
		return MC_ERR_STATE;
	*n = c->neta;
	return MC_OK;
}

#define NEED_MODEL() do { if (!c) return MC_ERR_ARG; if (!c->have_model) \
	return fail(c, MC_ERR_STATE, "%s: no model allocated", __func__); \
	CK(cudaSetDevice(c->device)); } while (0)
#define CHECK_SLOT(s) do { if ((s) < 0 || (s) > 2) \
	return fail(c, MC_ERR_ARG, "%s: slot %d out of range", __func__, (s)); } while (0)

extern "C" int mc_set_params(mc_ctx *c, int slot, const double *eta, const double *p)
{
	NEED_MODEL();
	CHECK_SLOT(slot);
	if (eta)
		CK(cudaMemcpyAsync(c->d_eta[slot], eta, sizeof(double) * (size_t)c->neta,
			cudaMemcpyHostToDevice, c->stream));
	if (p)
		CK(cudaMemcpyAsync(c->d_p[slot], p, sizeof(double) * (size_t)c->np,
			cudaMemcpyHostToDevice, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return MC_OK;
}

extern "C" int mc_get_params(mc_ctx *c, int slot, double *eta, double *p)
{
	NEED_MODEL();
	CHECK_SLOT(slot);
	if (eta)
		CK(cudaMemcpyAsync(eta, c->d_eta[slot], sizeof(double) * (size_t)c->neta,
			cudaMemcpyDeviceToHost, c->stream));
	if (p)
		CK(cudaMemcpyAsync(p, c->d_p[slot], sizeof(double) * (size_t)c->np,
			cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return MC_OK;
}

extern "C" int mc_copy_slot(mc_ctx *c, int dst, int src)
{
	NEED_MODEL();
	CHECK_SLOT(dst);
	CHECK_SLOT(src);
	if (dst == src)
		return MC_OK;
	CK(cudaMemcpyAsync(c->d_eta[dst], c->d_eta[src], sizeof(double) * (size_t)c->neta,
		cudaMemcpyDeviceToDevice, c->stream));
	CK(cudaMemcpyAsync(c->d_p[dst], c->d_p[src], sizeof(double) * (size_t)c->np,
		cudaMemcpyDeviceToDevice, c->stream));
	return MC_OK;
}

/* --------------------------------------------------- tile kernel dispatch */

static int launch_tile(mc_ctx *c, int mode, const double *p, const double *eta,
	long long eta_stride, bool fallback = false)
{
	tile_fn fn = mc_pick_tile(mode, c->KH, c->PP);
	if (!fn)
		return fail(c, MC_ERR_UNSUPPORTED, "no kernel for KH=%d PP=%d", c->KH, c->PP);
	TileArgs ta = c->ta;
	ta.p = p; ta.eta = eta; ta.eta_stride = eta_stride;
	if (fallback) {
		ta.run_if = c->d_dg_flag;
		ta.n_chunks_dev = c->d_dg_flag + 1;
	}
	{
		const int rca = raise_smem_limit(c, (const void *)fn, c->smem_em);
		if (rca)
			return rca;
	}
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	if (c->profile) {
		CK(cudaEventCreate(&e0));
		CK(cudaEventCreate(&e1));
		CK(cudaEventRecord(e0, c->stream));
	}
	fn<<<c->grid, c->block, c->smem_em, c->stream>>>(ta);
	LAUNCH_CHECK("tile_kernel");
	if (c->profile) {
		CK(cudaEventRecord(e1, c->stream));
		c->prof_events.push_back({ e0, e1 });
	}
	return MC_OK;
}

/* sum d_llpart[0..n) (or any vector) into *out on the device */
static int reduce_vector(mc_ctx *c, const double *x, long long n, double *out)
{
	if (n <= 16 * RED_THREADS) {	/* short: the final stage alone, same fixed order per n */
		k_colsum_final<<<1, RED_THREADS, 0, c->stream>>>(x, (int)n, 1, out);
		LAUNCH_CHECK("k_colsum_final");
		return MC_OK;
	}
	k_colsum_partial<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(x, n, 1, 1, c->d_red);
	LAUNCH_CHECK("k_colsum_partial");
	k_colsum_final<<<1, RED_THREADS, 0, c->stream>>>(c->d_red, RED_BLOCKS, 1, out);
	LAUNCH_CHECK("k_colsum_final");
	return MC_OK;
}

/* column sums of an [rows][K] matrix into out[0..K) */
static int reduce_columns(mc_ctx *c, const double *x, long long rows, int K, double *out)
{
	for (int k0 = 0; k0 < K; k0 += 8) {
		const int nc = std::min(8, K - k0);
		k_colsum_partial<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(x + k0, rows, nc, K, c->d_red);
		LAUNCH_CHECK("k_colsum_partial");
		k_colsum_final<<<1, RED_THREADS, 0, c->stream>>>(c->d_red, RED_BLOCKS, nc, out + k0);
		LAUNCH_CHECK("k_colsum_final");
	}
	return MC_OK;
}

/* -------------------------------------------------------------- EM step */

/* layout of the exchange buffer */
static inline double *xb_N(mc_ctx *c) { return c->d_xbuf; }
static inline double *xb_ll(mc_ctx *c) { return c->d_xbuf + c->np; }
static inline double *xb_S(mc_ctx *c) { return c->d_xbuf + c->np + 1; }

/* the main stream waits for the eta side of the last mc_em_step_local */
static int join_aux(mc_ctx *c)
{
	if (c->aux_pending) {
		c->aux_pending = false;
		CK(cudaStreamWaitEvent(c->stream, c->ev_aux, 0));
	}
	return MC_OK;
}

/* mixture E-step tail over the chunk partial sums in Apart: posterior rows (or
 * the guarded log-sum-exp of the log-likelihood pass), and -- fused in -- the
 * log likelihood and the pooled sums S_k = sum_i v_ik into the exchange buffer */
static int mix_tail(mc_ctx *c, const double *eta, double *vik, int ll_only)
{
	/* digit-sliced plan: the log-likelihood pass may have fallen back to the
	 * dense kernels, which cut the loci differently -- the kernel that ran left
	 * its chunk count on the device.  A table that cannot be represented in
	 * the E-step (a NaN in p) poisons the log likelihood instead. */
	const int *n_chunks_dev = c->use_dg && ll_only ? c->d_dg_flag + 1 : nullptr;
	int *flag = c->use_dg ? c->d_dg_flag : nullptr;
	const long long blocks = (c->I + MT_ROWS - 1) / MT_ROWS;
	const int grid = (int)std::min<long long>(blocks, (long long)c->num_sms * 4);
	/* per-block partial sums: grid * (2 K + 1) doubles */
	double *part = c->d_lli;
	if ((long long)grid * (2 * c->K + 1) > c->lli_n)
		return fail(c, MC_ERR_STATE, "mix_tail: partial-sum buffer too small");
	k_mix_tail<<<grid, 256, sizeof(double) * (MT_ROWS * c->K + MT_ROWS), c->stream>>>(c->d_Apart,
		c->act_tiles, n_chunks_dev, c->act_Ipad, c->I, c->K, eta, vik, part, ll_only);
	LAUNCH_CHECK("k_mix_tail");
	k_mix_final<<<1, 256, 0, c->stream>>>(part, grid, c->K, xb_ll(c), xb_S(c), ll_only, flag,
		c->use_dg && !ll_only ? c->d_dg_vscale : nullptr);
	LAUNCH_CHECK("k_mix_final");
	return MC_OK;
}

extern "C" int mc_em_step_local(mc_ctx *c, int from, int to)
{
	NVTX_FN();
	NEED_MODEL();
	CHECK_SLOT(from);
	CHECK_SLOT(to);
	const int K = c->K;
	int rc;
	if (c->admixture) {
		rc = c->use_dn
			? launch_dense(c, DN_ADMIX_EM, c->d_p[from], c->d_p[from], c->d_eta[from],
				c->per_indiv ? K : 0)
			: c->use3
			? launch_admix3(c, A3_ADMIX_EM, c->d_p[from], c->d_eta[from], c->per_indiv ? K : 0)
			: launch_tile(c, MODE_ADMIX_EM, c->d_p[from], c->d_eta[from],
				c->per_indiv ? K : 0);
		if (rc) return rc;
		if (!c->fused_step) {
			k_sum_chunks<<<grid_for(c, c->np, 256), 256, 0, c->stream>>>(c->d_Npart,
				c->act_chunks, c->np, 0.0, xb_N(c));
			LAUNCH_CHECK("k_sum_chunks");
		}
		if ((rc = reduce_vector(c, c->d_llpart, c->act_units, xb_ll(c)))) return rc;
		/* eta side: D_ik = eta_ik * A_ik needs only this context's
		 * individuals; the streaming kernel is done with slot `from`, so
		 * from == to (in-place EM) is safe */
		const int er = eta_rows(K);
		/* an individual-sharded fit exchanges the allele sums next; the eta
		 * side needs no remote data and runs beside the exchange (a pooled
		 * eta is part of the exchange and stays on the main stream) */
		cudaStream_t es = c->stream;
		if (c->per_indiv) {
			CK(cudaEventRecord(c->ev_main, c->stream));
			CK(cudaStreamWaitEvent(c->aux, c->ev_main, 0));
			es = c->aux;
		}
		k_admix_eta<<<(unsigned)std::min<long long>((c->I + er - 1) / er,
			(long long)c->num_sms * 16), 256, sizeof(double) * er * K, es>>>(c->d_Apart,
			c->act_tiles, c->act_Ipad, c->I, K, c->d_eta[from],
			c->per_indiv ? K : 0, c->d_eta[to], c->d_post, c->per_indiv,
			c->do_proj, c->eta_lb, er);
		LAUNCH_CHECK("k_admix_eta");
		if (c->per_indiv) {
			CK(cudaEventRecord(c->ev_aux, c->aux));
			c->aux_pending = true;
		} else {	/* pooled eta: S_k = sum_i D_ik */
			if ((rc = reduce_columns(c, c->d_post, c->I, K, xb_S(c)))) return rc;
		}
	} else {
		if (!c->use_dn && !c->use_dg) {
			k_log_table<<<grid_for(c, c->np, 256), 256, 0, c->stream>>>(c->d_p[from],
				c->d_logp, c->np, 1, nullptr);
			LAUNCH_CHECK("k_log_table");
		}
		if ((rc = c->use_dg ? launch_digit(c, DG_MIX_E, c->d_p[from], 1)
			: c->use_dn ? launch_dense(c, DN_MIX_E, c->d_p[from], nullptr, nullptr, 0, 1)
			: c->use3 ? launch_admix3(c, A3_MIX_E, c->d_logp, nullptr, 0)
			: launch_tile(c, MODE_MIX_E, c->d_logp, nullptr, 0))) return rc;
		if ((rc = mix_tail(c, c->d_eta[from], c->d_post, 0))) return rc;
		if ((rc = c->use_dg ? launch_digit(c, DG_MIX_M, c->d_post, 0)
			: c->use_dn ? launch_dense(c, DN_MIX_M, nullptr, nullptr, c->d_post, K)
			: c->use3 ? launch_admix3(c, A3_MIX_M, nullptr, c->d_post, K)
			: launch_tile(c, MODE_MIX_M, nullptr, c->d_post, K))) return rc;
		if (!c->fused_step) {
			k_sum_chunks<<<grid_for(c, c->np, 256), 256, 0, c->stream>>>(c->d_Npart,
				c->act_chunks, c->np, 0.0, xb_N(c));
			LAUNCH_CHECK("k_sum_chunks");
		}
		/* S_k = sum_i v_ik is already in the exchange buffer (mix_tail) */
	}
	return MC_OK;
}

extern "C" int mc_exchange_buffer(mc_ctx *c, void **dev_ptr, size_t *n)
{
	NEED_MODEL();
	if (dev_ptr) *dev_ptr = c->d_xbuf;
	if (n) *n = (size_t)c->np + 1 + c->K;
	return MC_OK;
}

extern "C" int mc_exchange_sum(mc_ctx *c, const void *gathered, int n_ranks)
{
	NEED_MODEL();
	if (!gathered || n_ranks < 1)
		return fail(c, MC_ERR_ARG, "mc_exchange_sum: bad arguments");
	const long long n = c->np + 1 + c->K;
	k_sum_chunks<<<grid_for(c, n, 256), 256, 0, c->stream>>>(
		(const double *)gathered, n_ranks, n, 0.0, c->d_xbuf);
	LAUNCH_CHECK("k_sum_chunks");
	return MC_OK;
}

extern "C" int mc_exchange_sum_slice(mc_ctx *c, const void *parts, int n_ranks,
	int64_t first, int64_t count)
{
	NVTX_FN();
	NEED_MODEL();
	const int64_t cap = c->np + 1 + c->K + 64;
	if (!parts || n_ranks < 1 || first < 0 || count < 0 || first + count > cap)
		return fail(c, MC_ERR_ARG, "mc_exchange_sum_slice: bad arguments");
	if (!count)
		return MC_OK;
	k_sum_chunks<<<grid_for(c, count, 256), 256, 0, c->stream>>>(
		(const double *)parts, n_ranks, count, 0.0, c->d_xbuf + first);
	LAUNCH_CHECK("k_sum_chunks");
	return MC_OK;
}

extern "C" int mc_em_step_finish(mc_ctx *c, int to, double *ll)
{
	NVTX_FN();
	NEED_MODEL();
	CHECK_SLOT(to);
	const int K = c->K;
	int rcj = join_aux(c);
	if (rcj)
		return rcj;
	/* the mixture adds the pseudo-count p_lower_bound to every slot
	 * (em_alg.c:972).  A whole step on one context (mc_em_step) reads the
	 * chunk sums directly; a split step reads the exchanged totals.  The pooled
	 * eta update rides along */
	const double add = c->admixture ? 0.0 : c->p_lb;
	const double *S = c->per_indiv ? nullptr : xb_S(c);
	if (c->fused_step)
		k_update_p<<<grid_for(c, (long long)K * c->L, 128), 128, 0, c->stream>>>(
			c->d_Npart, c->act_chunks, c->np, add, c->d_p[to], c->d_J, c->d_off, K,
			c->L, c->T, c->do_proj, c->p_lb, S, c->d_eta[to]);
	else
		k_update_p<<<grid_for(c, (long long)K * c->L, 128), 128, 0, c->stream>>>(
			xb_N(c), 1, 0, add, c->d_p[to], c->d_J, c->d_off, K, c->L, c->T,
			c->do_proj, c->p_lb, S, c->d_eta[to]);
	LAUNCH_CHECK("k_update_p");
	c->fused_step = false;
	if (ll) {
		CK(cudaMemcpyAsync(ll, xb_ll(c), sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
	}
	return MC_OK;
}

/* admixture initialiser: hard assignment + M-step (rnd_init.c:349-357) */
extern "C" int mc_init_admixture(mc_ctx *c, int slot, const uint8_t *z)
{
	int rc = mc_init_admixture_local(c, slot, z);
	if (rc)
		return rc;
	rc = mc_em_step_finish(c, slot, nullptr);
	CK(cudaStreamSynchronize(c->stream));
	return rc;
}

/* scratch buffer of the initialiser, grown when needed and kept between fits
 * (cudaMalloc / cudaFree per fit cost more than a small fit's EM) */
template <typename T> static int init_scratch(mc_ctx *c, T *&buf, size_t &have, size_t want)
{
	if (want <= have)
		return MC_OK;
	dfree(buf);
	have = 0;
	CK(MC_DEV_MALLOC(&buf, sizeof(T) * want));
	have = want;
	return MC_OK;
}

/* counts + M-step from the assignment in c->d_init_z */
static int init_from_assignment(mc_ctx *c, int slot)
{
	const size_t np = (size_t)std::max<int64_t>(c->np, 1);
	int rc = init_scratch(c, c->d_init_N, c->init_N_n, np);
	if (rc)
		return rc;
	CK(cudaMemsetAsync(c->d_init_N, 0, sizeof(unsigned) * np, c->stream));
	k_init_counts<<<(unsigned)std::min<long long>(std::max<long long>(c->I, 1),
		(long long)c->num_sms * 32), 128, 0, c->stream>>>(
		/* rnd_init.c:460-481 reads dat->IL, which a bootstrap never rewrites */
		c->d_nat_orig ? c->d_nat_orig : c->d_nat, c->d_init_z,
		c->I, c->L, c->P, c->K, c->d_off, c->T, c->d_post, c->d_init_N);
	LAUNCH_CHECK("k_init_counts");
	k_u32_to_f64<<<grid_for(c, c->np, 256), 256, 0, c->stream>>>(c->d_init_N, xb_N(c), c->np);
	LAUNCH_CHECK("k_u32_to_f64");
	if (c->per_indiv) {
		k_eta_from_D<<<grid_for(c, c->I, 128), 128, 0, c->stream>>>(c->d_post,
			c->d_eta[slot], c->I, c->K, c->do_proj, c->eta_lb);
		LAUNCH_CHECK("k_eta_from_D");
	} else {
		if ((rc = reduce_columns(c, c->d_post, c->I, c->K, xb_S(c)))) return rc;
	}
	return MC_OK;
}

static int init_checks(mc_ctx *c, int slot, const void *arg)
{
	NEED_MODEL();
	CHECK_SLOT(slot);
	if (!arg)
		return fail(c, MC_ERR_ARG, "mc_init_admixture: null argument");
	if (!c->admixture)
		return fail(c, MC_ERR_STATE, "mc_init_admixture: not an admixture model");
	if (c->K > 255)
		return fail(c, MC_ERR_UNSUPPORTED, "mc_init_admixture: K > 255");
	return MC_OK;
}

extern "C" int mc_init_admixture_local(mc_ctx *c, int slot, const uint8_t *z)
{
	NVTX_FN();
	int rc = init_checks(c, slot, z);
	if (rc)
		return rc;
	const size_t n = (size_t)c->I * c->L * c->P;
	if ((rc = init_scratch(c, c->d_init_z, c->init_z_n, n ? n : 1)))
		return rc;
	CK(cudaMemcpyAsync(c->d_init_z, z, n, cudaMemcpyHostToDevice, c->stream));
	rc = init_from_assignment(c, slot);
	CK(cudaStreamSynchronize(c->stream));	/* z is the caller's again */
	return rc;
}

extern "C" int mc_init_admixture_rand_local(mc_ctx *c, int slot, const uint32_t *hist,
	int64_t n_blocks, int64_t block_draws)
{
	NVTX_FN();
	int rc = init_checks(c, slot, hist);
	if (rc)
		return rc;
	const long long n = (long long)c->I * c->L * c->P;
	if (block_draws < 1 || block_draws % 16 || n_blocks * block_draws < n
		|| (n_blocks - 1) * block_draws >= std::max<long long>(n, 1))
		return fail(c, MC_ERR_ARG, "mc_init_admixture_rand: %lld blocks of %lld draws "
			"do not tile the %lld allele copies (blocks must be multiples of 16)",
			(long long)n_blocks, (long long)block_draws, n);
	if ((rc = init_scratch(c, c->d_init_z, c->init_z_n, n ? (size_t)n : 1))
		|| (rc = init_scratch(c, c->d_init_h, c->init_h_n, 31 * (size_t)n_blocks)))
		return rc;
	CK(cudaMemcpyAsync(c->d_init_h, hist, sizeof(unsigned) * 31 * (size_t)n_blocks,
		cudaMemcpyHostToDevice, c->stream));
	k_rand_assign<<<(unsigned)((n_blocks + 63) / 64), 64, 0, c->stream>>>(c->d_init_h,
		n_blocks, block_draws, n, (unsigned)c->K, c->d_init_z);
	LAUNCH_CHECK("k_rand_assign");
	rc = init_from_assignment(c, slot);
	CK(cudaStreamSynchronize(c->stream));	/* hist is the caller's again */
	return rc;
}

extern "C" int mc_init_admixture_rand(mc_ctx *c, int slot, const uint32_t *hist,
	int64_t n_blocks, int64_t block_draws)
{
	int rc = mc_init_admixture_rand_local(c, slot, hist, n_blocks, block_draws);
	if (rc)
		return rc;
	rc = mc_em_step_finish(c, slot, nullptr);
	CK(cudaStreamSynchronize(c->stream));
	return rc;
}

/* ------------------------------------------------- parametric bootstrap */

extern "C" int mc_save_mle(mc_ctx *c, int slot)
{
	NVTX_FN();
	NEED_MODEL();
	CHECK_SLOT(slot);
	const size_t ne = (size_t)(c->per_indiv ? c->I * (long long)c->K : c->K);
	const size_t np = (size_t)std::max<int64_t>(c->np, 1);
	if (c->mle_K != c->K || c->mle_per_indiv != c->per_indiv) {
		dfree(c->d_mle_eta); dfree(c->d_mle_p);
		CK(MC_DEV_MALLOC(&c->d_mle_eta, sizeof(double) * std::max<size_t>(ne, 1)));
		CK(MC_DEV_MALLOC(&c->d_mle_p, sizeof(double) * np));
		c->mle_K = c->K;
		c->mle_per_indiv = c->per_indiv;
	}
	c->mle_admixture = c->admixture;
	CK(cudaMemcpyAsync(c->d_mle_eta, c->d_eta[slot], sizeof(double) * ne,
		cudaMemcpyDeviceToDevice, c->stream));
	CK(cudaMemcpyAsync(c->d_mle_p, c->d_p[slot], sizeof(double) * (size_t)c->np,
		cudaMemcpyDeviceToDevice, c->stream));
	return MC_OK;
}

/* every data-derived layout goes; the model with them (its plan points there) */
static void drop_layouts(mc_ctx *c)
{
	free_model(c);
	free_layout3(c);
	free_layout_dg(c);
	c->dn_maxcode = -1;
}

extern "C" int mc_bootstrap_data(mc_ctx *c, const uint32_t *hist, int64_t n_blocks,
	int64_t block_draws)
{
	NVTX_FN();
	if (!c || !hist)
		return MC_ERR_ARG;
	if (!c->I || !c->d_mle_p)
		return fail(c, MC_ERR_STATE, "mc_bootstrap_data: no data or no saved estimates "
			"(mc_save_mle)");
	CK(cudaSetDevice(c->device));
	const long long per = (long long)c->L * c->P;
	const long long per_i = c->mle_admixture ? 2 * per : 1 + per;
	const long long n = c->I * per_i;
	if (block_draws < 1 || block_draws % 16 || n_blocks * block_draws < n
		|| (n_blocks - 1) * block_draws >= std::max<long long>(n, 1))
		return fail(c, MC_ERR_ARG, "mc_bootstrap_data: %lld blocks of %lld draws do not "
			"tile the %lld draws of the sample (blocks must be multiples of 16)",
			(long long)n_blocks, (long long)block_draws, n);
	drop_layouts(c);
	const size_t bytes = (size_t)c->I * c->L * c->P;
	if (!c->d_nat_orig) {
		c->d_nat_orig = c->d_nat;
		c->d_nat = nullptr;
		CK(MC_DEV_MALLOC(&c->d_nat, std::max<size_t>(bytes, 1)));
	}
	unsigned *d_hist = nullptr, *d_draws = nullptr;
	/* the scratch is released on every path */
	auto draw = [&]() -> int {
		CK(MC_DEV_MALLOC(&d_hist, sizeof(unsigned) * 31 * (size_t)n_blocks));
		CK(cudaMemcpyAsync(d_hist, hist, sizeof(unsigned) * 31 * (size_t)n_blocks,
			cudaMemcpyHostToDevice, c->stream));
		/* slabs of whole individuals, at most 2^28 draws (1 GiB of raw draws) each */
		const long long slab_i = std::max<long long>(1, (1LL << 28) / per_i);
		const long long cap_blocks = (slab_i * per_i + block_draws - 1) / block_draws + 2;
		CK(MC_DEV_MALLOC(&d_draws, sizeof(unsigned) * (size_t)std::min<long long>(cap_blocks,
			n_blocks) * block_draws));
		for (long long i0 = 0; i0 < c->I; i0 += slab_i) {
			const long long i1 = std::min<long long>(c->I, i0 + slab_i);
			const long long b0 = i0 * per_i / block_draws;
			const long long b1 = (i1 * per_i - 1) / block_draws + 1;
			const long long d0 = b0 * block_draws;
			k_rand_raw<<<(unsigned)((b1 - b0 + 63) / 64), 64, 0, c->stream>>>(
				d_hist + 31 * b0, b1 - b0, block_draws, n - d0, d_draws);
			LAUNCH_CHECK("k_rand_raw");
			k_bootstrap_codes<<<grid_for(c, (i1 - i0) * c->L, 128), 128, 0, c->stream>>>(
				d_draws, d0, i0, i1, c->L, c->P, c->mle_K, c->T, c->d_off, c->d_J,
				c->d_mle_eta, c->mle_per_indiv ? c->mle_K : 0, c->d_mle_p,
				c->mle_admixture, c->d_nat);
			LAUNCH_CHECK("k_bootstrap_codes");
		}
		CK(cudaStreamSynchronize(c->stream));	/* hist is the caller's again */
		return MC_OK;
	};
	const int rc = draw();
	cudaFree(d_hist);
	cudaFree(d_draws);
	return rc;
}

extern "C" int mc_restore_data(mc_ctx *c)
{
	if (!c)
		return MC_ERR_ARG;
	if (!c->d_nat_orig)
		return MC_OK;
	CK(cudaSetDevice(c->device));
	drop_layouts(c);
	dfree(c->d_nat);
	c->d_nat = c->d_nat_orig;
	c->d_nat_orig = nullptr;
	return MC_OK;
}

/* mixture initialiser (rnd_init.c:192-339): nearest-centre assignment of this
 * context's individuals and the cluster's allele counts, left in the exchange
 * buffer [K*T counts | - | K cluster sizes] */
extern "C" int mc_init_mixture_local(mc_ctx *c, const int32_t *center_idx,
	const uint8_t *center_codes)
{
	NVTX_FN();
	NEED_MODEL();
	if (c->admixture)
		return fail(c, MC_ERR_STATE, "mc_init_mixture: not a mixture model");
	if (c->K > 1 && (!center_idx || !center_codes))
		return fail(c, MC_ERR_ARG, "mc_init_mixture: null argument");
	const int K = c->K;
	const size_t np = (size_t)std::max<int64_t>(c->np, 1), row = (size_t)c->L * c->P;
	int rc = init_scratch(c, c->d_init_N, c->init_N_n, np);
	if (rc)
		return rc;
	unsigned char *d_cent = nullptr;
	int *d_cidx = nullptr;
	unsigned *d_nk = nullptr;
	CK(MC_DEV_MALLOC(&d_cent, std::max<size_t>(row * K, 1)));
	CK(MC_DEV_MALLOC(&d_cidx, sizeof(int) * K));
	CK(MC_DEV_MALLOC(&d_nk, sizeof(unsigned) * K));
	std::vector<int> none((size_t)K, -1);
	CK(cudaMemcpyAsync(d_cidx, K > 1 ? center_idx : none.data(), sizeof(int) * K,
		cudaMemcpyHostToDevice, c->stream));
	if (K > 1)
		CK(cudaMemcpyAsync(d_cent, center_codes, row * K, cudaMemcpyHostToDevice, c->stream));
	CK(cudaMemsetAsync(d_nk, 0, sizeof(unsigned) * K, c->stream));
	CK(cudaMemsetAsync(c->d_init_N, 0, sizeof(unsigned) * np, c->stream));
	const unsigned grid = (unsigned)std::min<long long>(std::max<long long>(c->I, 1),
		(long long)c->num_sms * 32);
	k_mix_assign<<<grid, 128, 0, c->stream>>>(c->d_nat, d_cent, d_cidx, c->I, c->L, c->P, K,
		c->d_IK, d_nk);
	LAUNCH_CHECK("k_mix_assign");
	k_mix_init_counts<<<grid, 128, 0, c->stream>>>(c->d_nat, c->d_IK, c->I, c->L, c->P,
		c->d_off, c->T, c->d_init_N);
	LAUNCH_CHECK("k_mix_init_counts");
	k_u32_to_f64<<<grid_for(c, c->np, 256), 256, 0, c->stream>>>(c->d_init_N, xb_N(c), c->np);
	LAUNCH_CHECK("k_u32_to_f64");
	k_u32_to_f64<<<1, 32, 0, c->stream>>>(d_nk, xb_S(c), K);
	LAUNCH_CHECK("k_u32_to_f64");
	CK(cudaStreamSynchronize(c->stream));	/* the host arrays are the caller's again */
	cudaFree(d_cent); cudaFree(d_cidx); cudaFree(d_nk);
	return MC_OK;
}

extern "C" int mc_init_mixture_finish(mc_ctx *c, int slot, int64_t I_total)
{
	NEED_MODEL();
	CHECK_SLOT(slot);
	if (c->admixture)
		return fail(c, MC_ERR_STATE, "mc_init_mixture: not a mixture model");
	k_mix_init_finish<<<grid_for(c, (long long)c->K * c->L, 128), 128, 0, c->stream>>>(xb_N(c),
		xb_S(c), c->d_eta[slot], c->d_p[slot], c->d_J, c->d_off, c->K, c->L, c->T, I_total);
	LAUNCH_CHECK("k_mix_init_finish");
	return MC_OK;
}

extern "C" int mc_init_mixture(mc_ctx *c, int slot, const int32_t *center_idx,
	const uint8_t *center_codes)
{
	int rc = mc_init_mixture_local(c, center_idx, center_codes);
	if (rc)
		return rc;
	rc = mc_init_mixture_finish(c, slot, c->I);
	CK(cudaStreamSynchronize(c->stream));
	return rc;
}

/* Capture `body` (kernel launches on c->stream and c->aux only) into a graph.
 * Returns MC_OK with *exec set, or MC_OK with *exec null when capture is not
 * possible here (the caller then runs the body directly, and graphs are
 * switched off for this plan). */
template <typename F>
static int capture_graph(mc_ctx *c, cudaGraphExec_t *exec, int64_t *n_kernels, F body)
{
	*exec = nullptr;
	if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
		(void)cudaGetLastError();
		c->graphs_ok = false;
		return MC_OK;
	}
	const int64_t l0 = c->launches;
	const int rc = body();
	cudaGraph_t graph = nullptr;
	const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
	*n_kernels = c->launches - l0;
	c->launches = l0;
	if (rc || e != cudaSuccess || !graph
		|| cudaGraphInstantiate(exec, graph, 0) != cudaSuccess) {
		(void)cudaGetLastError();
		if (graph)
			cudaGraphDestroy(graph);
		*exec = nullptr;
		c->graphs_ok = false;
		c->aux_pending = false;
		c->fused_step = false;
		return rc;
	}
	cudaGraphDestroy(graph);
	return MC_OK;
}

static int read_ll(mc_ctx *c, double *ll)
{
	if (ll) {
		CK(cudaMemcpyAsync(ll, xb_ll(c), sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
	}
	return MC_OK;
}

extern "C" int mc_em_step(mc_ctx *c, int from, int to, double *ll)
{
	if (!c)
		return MC_ERR_ARG;
	if (!c->have_model)
		return fail(c, MC_ERR_STATE, "mc_em_step: no model allocated");
	CHECK_SLOT(from);
	CHECK_SLOT(to);
	auto body = [&]() {
		c->fused_step = true;	/* nobody exchanges the allele sums in between */
		int rc = mc_em_step_local(c, from, to);
		if (rc) {
			c->fused_step = false;
			return rc;
		}
		return mc_em_step_finish(c, to, nullptr);
	};
	/* the first call of a slot pair runs directly (it also raises the kernels'
	 * shared-memory limits), the second one is captured, later ones replayed */
	const int key = from * 3 + to;
	if (c->opt_graph && c->graphs_ok && !c->profile) {
		CK(cudaSetDevice(c->device));
		if (!c->g_step[key] && c->g_step_seen[key]++ >= 1) {
			const int rc = capture_graph(c, &c->g_step[key], &c->g_step_n[key], body);
			if (rc)
				return rc;
		}
		if (c->g_step[key]) {
			CK(cudaGraphLaunch(c->g_step[key], c->stream));
			c->launches += c->g_step_n[key];
			return read_ll(c, ll);
		}
	}
	const int rc = body();
	return rc ? rc : read_ll(c, ll);
}

static int loglik_launch(mc_ctx *c, int slot);

extern "C" int mc_loglik(mc_ctx *c, int slot, double *ll)
{
	NVTX_FN();
	NEED_MODEL();
	CHECK_SLOT(slot);
	if (c->opt_graph && c->graphs_ok && !c->profile) {
		if (!c->g_ll[slot] && c->g_ll_seen[slot]++ >= 1) {
			const int rc = capture_graph(c, &c->g_ll[slot], &c->g_ll_n[slot],
				[&]() { return loglik_launch(c, slot); });
			if (rc)
				return rc;
		}
		if (c->g_ll[slot]) {
			CK(cudaGraphLaunch(c->g_ll[slot], c->stream));
			c->launches += c->g_ll_n[slot];
			return read_ll(c, ll);
		}
	}
	const int rc = loglik_launch(c, slot);
	return rc ? rc : read_ll(c, ll);
}

static int loglik_launch(mc_ctx *c, int slot)
{
	int rc;
	if (c->admixture) {
		rc = c->use_dn
			? launch_dense(c, DN_ADMIX_LL, c->d_p[slot], c->d_p[slot], c->d_eta[slot],
				c->per_indiv ? c->K : 0)
			: c->use3
			? launch_admix3(c, A3_ADMIX_LL, c->d_p[slot], c->d_eta[slot], c->per_indiv ? c->K : 0)
			: launch_tile(c, MODE_ADMIX_LL, c->d_p[slot], c->d_eta[slot],
				c->per_indiv ? c->K : 0);
		if (rc) return rc;
		if ((rc = reduce_vector(c, c->d_llpart, c->act_units, xb_ll(c)))) return rc;
	} else {
		/* the digit-sliced pass first (it raises the flag), then the FP64 pass:
		 * the pass proper, or the fall-back gated on the flag */
		const bool fb = c->use_dg;
		if (fb && (rc = launch_digit(c, DG_MIX_E, c->d_p[slot], 2))) return rc;
		if (!c->use_dn) {
			k_log_table<<<grid_for(c, c->np, 256), 256, 0, c->stream>>>(c->d_p[slot],
				c->d_logp, c->np, 0, fb ? c->d_dg_flag : nullptr);
			LAUNCH_CHECK("k_log_table");
		}
		if ((rc = c->use_dn ? launch_dense(c, DN_MIX_E, c->d_p[slot], nullptr, nullptr, 0, 2, fb)
			: c->use3 ? launch_admix3(c, A3_MIX_E, c->d_logp, nullptr, 0, fb)
			: launch_tile(c, MODE_MIX_E, c->d_logp, nullptr, 0, fb))) return rc;
		/* the posterior of the last E-step must survive: only the per-
		 * individual ll buffer is written */
		if ((rc = mix_tail(c, c->d_eta[slot], nullptr, 1))) return rc;
	}
	return MC_OK;
}

extern "C" int mc_read_ll(mc_ctx *c, double *ll)
{
	NEED_MODEL();
	if (!ll)
		return fail(c, MC_ERR_ARG, "mc_read_ll: null pointer");
	CK(cudaMemcpyAsync(ll, xb_ll(c), sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return MC_OK;
}

extern "C" int mc_get_posterior(mc_ctx *c, double *out)
{
	NEED_MODEL();
	CK(cudaMemcpyAsync(out, c->d_post, sizeof(double) * (size_t)c->I * c->K,
		cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return MC_OK;
}

extern "C" int mc_locale_sums(mc_ctx *c, const int32_t *locale, int32_t n_locales, double *out)
{
	NEED_MODEL();
	if (!locale || !out || n_locales < 1)
		return fail(c, MC_ERR_ARG, "mc_locale_sums: bad arguments");
	const int n = n_locales * c->K;
	const int blocks = (int)((c->I + LS_ROWS - 1) / LS_ROWS);
	int *d_loc = nullptr;
	double *d_part = nullptr, *d_out = nullptr;
	CK(MC_DEV_MALLOC(&d_loc, sizeof(int) * (size_t)c->I));
	CK(MC_DEV_MALLOC(&d_part, sizeof(double) * (size_t)blocks * n));
	CK(MC_DEV_MALLOC(&d_out, sizeof(double) * (size_t)n));
	CK(cudaMemcpyAsync(d_loc, locale, sizeof(int) * (size_t)c->I, cudaMemcpyHostToDevice, c->stream));
	k_locale_partial<<<blocks, 256, 0, c->stream>>>(c->d_post, d_loc, c->I, c->K, n_locales, d_part);
	LAUNCH_CHECK("k_locale_partial");
	k_locale_final<<<(n + 127) / 128, 128, 0, c->stream>>>(d_part, blocks, n, d_out);
	LAUNCH_CHECK("k_locale_final");
	CK(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	cudaFree(d_loc); cudaFree(d_part); cudaFree(d_out);
	return MC_OK;
}

extern "C" int mc_partition(mc_ctx *c, int32_t *I_K, int32_t *count_K)
{
	NEED_MODEL();
	k_partition<<<grid_for(c, c->I, 256), 256, 0, c->stream>>>(c->d_post, c->I, c->K, c->d_IK);
	LAUNCH_CHECK("k_partition");
	std::vector<int32_t> tmp;
	int32_t *dst = I_K;
	if (!dst) {
		tmp.resize((size_t)c->I);
		dst = tmp.data();
	}
	CK(cudaMemcpyAsync(dst, c->d_IK, sizeof(int) * (size_t)c->I, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	if (count_K) {
		for (int k = 0; k < c->K; k++)
			count_K[k] = 0;
		for (int64_t i = 0; i < c->I; i++)
			count_K[dst[i]]++;
	}
	return MC_OK;
}

/* ------------------------------------------------ acceleration plumbing */

#define CHECK_PAIR(x) do { if ((x) < 0 || (x) >= c->q) \
	return fail(c, MC_ERR_ARG, "%s: secant pair %d out of range (q=%d)", \
		__func__, (x), c->q); } while (0)

extern "C" int mc_delta(mc_ctx *c, int which, int pair, int slot_t, int slot_f)
{
	NVTX_FN();
	NEED_MODEL();
	CHECK_PAIR(pair);
	CHECK_SLOT(slot_t);
	CHECK_SLOT(slot_f);
	double *dp = which ? c->d_vp[pair] : c->d_up[pair];
	double *de = which ? c->d_ve[pair] : c->d_ue[pair];
	k_delta<<<grid_for(c, c->np, 256), 256, 0, c->stream>>>(dp, c->d_p[slot_t], c->d_p[slot_f], c->np);
	LAUNCH_CHECK("k_delta");
	k_delta<<<grid_for(c, c->neta, 256), 256, 0, c->stream>>>(de, c->d_eta[slot_t], c->d_eta[slot_f], c->neta);
	LAUNCH_CHECK("k_delta");
	return MC_OK;
}

static int fetch_small(mc_ctx *c, double *dst, int off, int n)
{
	CK(cudaMemcpyAsync(dst, c->d_small + off, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
	return MC_OK;
}

extern "C" int mc_step_dots(mc_ctx *c, int pair, double eta_part[3], double p_part[3])
{
	NVTX_FN();
	NEED_MODEL();
	CHECK_PAIR(pair);
	double z[3] = { 0, 0, 0 };
	int rc;
	k_step_dots<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_ue[pair], c->d_ve[pair], c->neta, c->d_red);
	LAUNCH_CHECK("k_step_dots");
	k_colsum_final<<<1, RED_THREADS, 0, c->stream>>>(c->d_red, RED_BLOCKS, 3, c->d_small);
	LAUNCH_CHECK("k_colsum_final");
	k_step_dots<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_up[pair], c->d_vp[pair], c->np, c->d_red + RED_BLOCKS * 3);
	LAUNCH_CHECK("k_step_dots");
	k_colsum_final<<<1, RED_THREADS, 0, c->stream>>>(c->d_red + RED_BLOCKS * 3, RED_BLOCKS, 3, c->d_small + 3);
	LAUNCH_CHECK("k_colsum_final");
	if ((rc = fetch_small(c, eta_part ? eta_part : z, 0, 3))) return rc;
	double z2[3];
	if ((rc = fetch_small(c, p_part ? p_part : z2, 3, 3))) return rc;
	CK(cudaStreamSynchronize(c->stream));
	return MC_OK;
}

extern "C" int mc_qn_dots(mc_ctx *c, int q1, int q2, double eta_part[2], double p_part[2])
{
	NVTX_FN();
	NEED_MODEL();
	CHECK_PAIR(q1);
	CHECK_PAIR(q2);
	double z[2], z2[2];
	int rc;
	k_qn_dots<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_ue[q1], c->d_ue[q2], c->d_ve[q2], c->neta, c->d_red);
	LAUNCH_CHECK("k_qn_dots");
	k_colsum_final<<<1, RED_THREADS, 0, c->stream>>>(c->d_red, RED_BLOCKS, 2, c->d_small);
	LAUNCH_CHECK("k_colsum_final");
	k_qn_dots<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_up[q1], c->d_up[q2], c->d_vp[q2], c->np, c->d_red + RED_BLOCKS * 3);
	LAUNCH_CHECK("k_qn_dots");
	k_colsum_final<<<1, RED_THREADS, 0, c->stream>>>(c->d_red + RED_BLOCKS * 3, RED_BLOCKS, 2, c->d_small + 3);
	LAUNCH_CHECK("k_colsum_final");
	if ((rc = fetch_small(c, eta_part ? eta_part : z, 0, 2))) return rc;
	if ((rc = fetch_small(c, p_part ? p_part : z2, 3, 2))) return rc;
	CK(cudaStreamSynchronize(c->stream));
	return MC_OK;
}

static int project_slot(mc_ctx *c, int slot)
{
	k_project_p<<<grid_for(c, (long long)c->K * c->L, 128), 128, 0, c->stream>>>(
		c->d_p[slot], c->d_J, c->d_off, c->K, c->L, c->T, c->p_lb);
	LAUNCH_CHECK("k_project_p");
	const long long rows = c->per_indiv ? c->I : 1;
	k_project_eta<<<grid_for(c, rows, 128), 128, 0, c->stream>>>(c->d_eta[slot], rows, c->K, c->eta_lb);
	LAUNCH_CHECK("k_project_eta");
	return MC_OK;
}

extern "C" int mc_project(mc_ctx *c, int slot)
{
	NEED_MODEL();
	CHECK_SLOT(slot);
	return project_slot(c, slot);
}

extern "C" int mc_accel_update(mc_ctx *c, int qn1, int slot_t, int slot_p, int pair, double s)
{
	NVTX_FN();
	NEED_MODEL();
	CHECK_PAIR(pair);
	CHECK_SLOT(slot_t);
	CHECK_SLOT(slot_p);
	k_accel_update<<<grid_for(c, c->np, 256), 256, 0, c->stream>>>(c->d_p[slot_t],
		c->d_p[slot_p], c->d_up[pair], c->d_vp[pair], c->np, s, qn1);
	LAUNCH_CHECK("k_accel_update");
	k_accel_update<<<grid_for(c, c->neta, 256), 256, 0, c->stream>>>(c->d_eta[slot_t],
		c->d_eta[slot_p], c->d_ue[pair], c->d_ve[pair], c->neta, s, qn1);
	LAUNCH_CHECK("k_accel_update");
	if (c->do_proj)
		return project_slot(c, slot_t);
	return MC_OK;
}

extern "C" int mc_qn_update(mc_ctx *c, int slot_t, int slot_p, int uindex,
	int delta_index, const double *Ainv, const double *cutu)
{
	NVTX_FN();
	NEED_MODEL();
	CHECK_SLOT(slot_t);
	CHECK_SLOT(slot_p);
	CHECK_PAIR(uindex);
	CHECK_PAIR(delta_index);
	if (!Ainv || !cutu)
		return fail(c, MC_ERR_ARG, "mc_qn_update: null coefficients");
	const int q = c->q;
	QnArgs qp, qe;
	memset(&qp, 0, sizeof qp);
	qp.q = q;
	/* row j uses the pair (delta_index + j) % q (accel_em.c:378-402) */
	for (int j = 0; j < q; j++)
		for (int n = 0; n < q; n++) {
			qp.coef[j][n][0] = Ainv[j * q + n];
			qp.coef[j][n][1] = cutu[n];
		}
	qe = qp;
	for (int j = 0; j < q; j++) {
		qp.v[j] = c->d_vp[(delta_index + j) % q];
		qe.v[j] = c->d_ve[(delta_index + j) % q];
	}
	k_qn_update<<<grid_for(c, c->np, 256), 256, 0, c->stream>>>(c->d_p[slot_t],
		c->d_p[slot_p], c->d_up[uindex], c->np, qp);
	LAUNCH_CHECK("k_qn_update");
	k_qn_update<<<grid_for(c, c->neta, 256), 256, 0, c->stream>>>(c->d_eta[slot_t],
		c->d_eta[slot_p], c->d_ue[uindex], c->neta, qe);
	LAUNCH_CHECK("k_qn_update");
	if (c->do_proj)
		return project_slot(c, slot_t);
	return MC_OK;
}

/* -------------------------------------------------------- introspection */

extern "C" int mc_get_plan(const mc_ctx *c, mc_plan_info *o)
{
	if (!c || !c->have_model || !o)
		return MC_ERR_STATE;
	o->K = c->K; o->k_split = c->k_split; o->k_per_lane = c->KH;
	o->loci_per_warp = c->ta.loci_per_warp; o->warps = c->ta.warps;
	o->groups = c->ta.groups; o->n_tiles = c->ta.n_tiles;
	o->n_chunks = c->ta.n_chunks; o->n_units = c->ta.n_units;
	o->grid = c->grid; o->block = c->block;
	o->indiv_per_block = c->IB; o->ploidy_padded = c->PP;
	o->smem_bytes = (int64_t)c->smem_em;
	o->two_pass = 0;
	if (c->use3) {	/* two-pass admixture kernel: tiles are locus chunks */
		o->two_pass = 2;
		o->k_split = 1; o->k_per_lane = 2 * c->KP3;
		o->loci_per_warp = A3_NC / c->PP; o->warps = A3_THREADS / 32; o->groups = 1;
		o->n_tiles = c->a3.n_lchunks; o->n_chunks = c->a3.n_ichunks;
		o->n_units = c->a3.n_units; o->grid = c->grid3; o->block = A3_THREADS;
		o->indiv_per_block = A3_IT; o->smem_bytes = (int64_t)c->smem3;
	}
	if (c->use_dn) {	/* dense DMMA kernels: tiles are locus chunks */
		o->two_pass = 3;
		o->k_split = 1; o->k_per_lane = 8 * c->dn_NB;
		o->loci_per_warp = DN_TL; o->warps = DN_THREADS / 32; o->groups = 1;
		o->n_tiles = c->dn.n_lchunks; o->n_chunks = c->dn.n_ichunks;
		o->n_units = c->dn.n_units; o->grid = c->grid_dn; o->block = DN_THREADS;
		o->indiv_per_block = DN_IT;
		o->smem_bytes = (int64_t)dn_smem_bytes(c->dn_NB,
			c->admixture ? DN_ADMIX_EM : DN_MIX_M, c->dn.max_chunk_tiles);
	}
	if (c->use_dg) {	/* digit-sliced mixture kernels: chunks of the two passes */
		o->two_pass = 4;
		o->k_per_lane = c->K;
		o->loci_per_warp = 64; o->warps = DG_WARPS;
		o->n_tiles = c->dgE.n_chunks; o->n_chunks = c->dgM.n_chunks;
		o->n_units = c->dgE.n_chunks * c->dgE.n_ctarows;
		o->grid = o->n_units; o->block = DG_THREADS;
		o->indiv_per_block = 16 * dg_R(c->K) * DG_WARPS;
		o->smem_bytes = (int64_t)dg_smem_bytes(c->K);
	}
	const int64_t g = c->I * (int64_t)c->L * c->P;
	o->algorithmic_bytes_em = g + 16 * c->I * (int64_t)c->K + 16 * (int64_t)c->K * c->T;
	o->algorithmic_bytes_ll = g + 8 * c->I * (int64_t)c->K + 8 * (int64_t)c->K * c->T;
	return MC_OK;
}

extern "C" int64_t mc_launch_count(const mc_ctx *c) { return c ? c->launches : 0; }

extern "C" int mc_profile_enable(mc_ctx *c, int on)
{
	if (!c)
		return MC_ERR_ARG;
	c->profile = on != 0;
	return MC_OK;
}

extern "C" int mc_profile_read(mc_ctx *c, int64_t *n, double *ms)
{
	if (!c)
		return MC_ERR_ARG;
	CK(cudaSetDevice(c->device));
	CK(cudaStreamSynchronize(c->stream));
	for (auto &ev : c->prof_events) {
		float t = 0;
		CK(cudaEventElapsedTime(&t, ev.first, ev.second));
		c->prof_ms += t;
		c->prof_n++;
		cudaEventDestroy(ev.first);
		cudaEventDestroy(ev.second);
	}
	c->prof_events.clear();
	if (n) *n = c->prof_n;
	if (ms) *ms = c->prof_ms;
	c->prof_n = 0;
	c->prof_ms = 0;
	return MC_OK;
}

extern "C" int mc_set_option(mc_ctx *c, int option, int value)
{
	if (!c)
		return MC_ERR_ARG;
	switch (option) {
	case MC_OPT_KERNEL:
		if (value < MC_KERNEL_AUTO || value > MC_KERNEL_DIGIT)
			return fail(c, MC_ERR_ARG, "mc_set_option: kernel %d out of range", value);
		c->opt_kernel = value;
		return MC_OK;
	case MC_OPT_TIMING:
		c->opt_timing = value != 0;
		return MC_OK;
	case MC_OPT_GRAPH:
		c->opt_graph = value != 0;
		return MC_OK;
	}
	return fail(c, MC_ERR_ARG, "mc_set_option: unknown option %d", option);
}
