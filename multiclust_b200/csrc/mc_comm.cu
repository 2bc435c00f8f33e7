/*
 * mc_comm.cu -- see include/mc_comm.h.
 * Build: nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC
 *        -Iinclude -o libmc_comm.so mc_comm.cu -L. -lmc_cuda -lnccl
 */
#include <cstdio>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>

#include "mc_comm.h"

struct mc_comm {
	int n = 0;
	std::vector<mc_ctx *> ctx;
	std::vector<int> dev;
	std::vector<ncclComm_t> nccl;
	std::vector<double *> gathered;		/* [n * n_doubles] on each device */
	size_t n_doubles = 0;
	std::string err;
};

static std::string g_err;

extern "C" const char *mc_comm_last_error(const mc_comm *c)
{
	return c ? c->err.c_str() : g_err.c_str();
}

extern "C" int mc_ctx_device(const mc_ctx *ctx);
extern "C" void *mc_ctx_stream(const mc_ctx *ctx);

extern "C" int mc_comm_create(mc_comm **out, mc_ctx **ctxs, int n)
{
	if (!out || !ctxs || n < 1) {
		g_err = "mc_comm_create: bad arguments";
		return MC_ERR_ARG;
	}
	mc_comm *c = new mc_comm();
	c->n = n;
	for (int r = 0; r < n; r++) {
		void *ptr = nullptr;
		size_t nd = 0;
		if (mc_exchange_buffer(ctxs[r], &ptr, &nd) != MC_OK || (r && nd != c->n_doubles)) {
			g_err = "mc_comm_create: contexts without a model, or of different sizes";
			delete c;
			return MC_ERR_STATE;
		}
		c->n_doubles = nd;
		c->ctx.push_back(ctxs[r]);
		c->dev.push_back(mc_ctx_device(ctxs[r]));
	}
	c->nccl.resize((size_t)n);
	ncclResult_t nr = ncclCommInitAll(c->nccl.data(), n, c->dev.data());
	if (nr != ncclSuccess) {
		g_err = std::string("ncclCommInitAll failed: ") + ncclGetErrorString(nr);
		delete c;
		return MC_ERR_CUDA;
	}
	for (int r = 0; r < n; r++) {
		double *g = nullptr;
		cudaSetDevice(c->dev[r]);
		if (cudaMalloc(&g, sizeof(double) * (c->n_doubles * n + 64)) != cudaSuccess) {
			g_err = "mc_comm_create: cudaMalloc failed";
			mc_comm_destroy(c);
			return MC_ERR_NOMEM;
		}
		c->gathered.push_back(g);
	}
	*out = c;
	return MC_OK;
}

extern "C" void mc_comm_destroy(mc_comm *c)
{
	if (!c)
		return;
	for (size_t r = 0; r < c->gathered.size(); r++) {
		cudaSetDevice(c->dev[r]);
		cudaFree(c->gathered[r]);
	}
	for (size_t r = 0; r < c->nccl.size(); r++)
		if (c->nccl[r])
			ncclCommDestroy(c->nccl[r]);
	delete c;
}

/* every NCCL call of one grouped operation; the group is always closed, the
 * first error is the one reported */
#define NCCL_IN_GROUP(call) do { ncclResult_t r_ = (call); \
	if (r_ != ncclSuccess && first == ncclSuccess) first = r_; } while (0)

extern "C" int mc_comm_exchange(mc_comm *c)
{
	if (!c)
		return MC_ERR_ARG;
	const int n = c->n;
	std::vector<double *> x((size_t)n);
	for (int r = 0; r < n; r++) {
		void *ptr = nullptr;
		const int rc = mc_exchange_buffer(c->ctx[r], &ptr, nullptr);
		if (rc != MC_OK) {
			c->err = mc_last_error(c->ctx[r]);
			return rc;
		}
		x[r] = static_cast<double *>(ptr);
	}
	ncclResult_t first = ncclSuccess;
	if (n <= 64) {
		/* reduce-scatter + all-gather in rank order (include/mc_cuda.h,
		 * mc_exchange_sum_slice): slice j of every buffer goes to device j */
		const size_t m = (c->n_doubles + n - 1) / n;
		NCCL_IN_GROUP(ncclGroupStart());
		for (int r = 0; r < n; r++) {
			cudaSetDevice(c->dev[r]);
			cudaStream_t st = (cudaStream_t)mc_ctx_stream(c->ctx[r]);
			for (int j = 0; j < n; j++) {
				NCCL_IN_GROUP(ncclSend(x[r] + j * m, m, ncclDouble, j, c->nccl[r], st));
				NCCL_IN_GROUP(ncclRecv(c->gathered[r] + j * m, m, ncclDouble, j, c->nccl[r], st));
			}
		}
		NCCL_IN_GROUP(ncclGroupEnd());
		if (first != ncclSuccess) {
			c->err = std::string("NCCL all-to-all failed: ") + ncclGetErrorString(first);
			return MC_ERR_CUDA;
		}
		for (int r = 0; r < n; r++) {
			const int rc = mc_exchange_sum_slice(c->ctx[r], c->gathered[r], n,
				(int64_t)(r * m), (int64_t)m);
			if (rc) {
				c->err = mc_last_error(c->ctx[r]);
				return rc;
			}
		}
		NCCL_IN_GROUP(ncclGroupStart());
		for (int r = 0; r < n; r++) {
			cudaSetDevice(c->dev[r]);
			/* in place: the send buffer is the rank's own slice of the result */
			NCCL_IN_GROUP(ncclAllGather(x[r] + r * m, x[r], m, ncclDouble, c->nccl[r],
				(cudaStream_t)mc_ctx_stream(c->ctx[r])));
		}
		NCCL_IN_GROUP(ncclGroupEnd());
		if (first != ncclSuccess) {
			c->err = std::string("ncclAllGather failed: ") + ncclGetErrorString(first);
			return MC_ERR_CUDA;
		}
		return MC_OK;
	}
	NCCL_IN_GROUP(ncclGroupStart());
	for (int r = 0; r < n; r++) {
		cudaSetDevice(c->dev[r]);
		NCCL_IN_GROUP(ncclAllGather(x[r], c->gathered[r], c->n_doubles, ncclDouble,
			c->nccl[r], (cudaStream_t)mc_ctx_stream(c->ctx[r])));
	}
	NCCL_IN_GROUP(ncclGroupEnd());
	if (first != ncclSuccess) {
		c->err = std::string("ncclAllGather failed: ") + ncclGetErrorString(first);
		return MC_ERR_CUDA;
	}
	for (int r = 0; r < n; r++) {
		const int rc = mc_exchange_sum(c->ctx[r], c->gathered[r], n);
		if (rc) {
			c->err = mc_last_error(c->ctx[r]);
			return rc;
		}
	}
	return MC_OK;
}
