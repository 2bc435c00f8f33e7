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
		if (cudaMalloc(&g, sizeof(double) * c->n_doubles * n) != cudaSuccess) {
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

extern "C" int mc_comm_exchange(mc_comm *c)
{
	if (!c)
		return MC_ERR_ARG;
	ncclResult_t nr = ncclGroupStart();
	for (int r = 0; r < c->n && nr == ncclSuccess; r++) {
		void *send = nullptr;
		mc_exchange_buffer(c->ctx[r], &send, nullptr);
		nr = ncclAllGather(send, c->gathered[r], c->n_doubles, ncclDouble,
			c->nccl[r], (cudaStream_t)mc_ctx_stream(c->ctx[r]));
	}
	if (nr == ncclSuccess)
		nr = ncclGroupEnd();
	if (nr != ncclSuccess) {
		c->err = std::string("ncclAllGather failed: ") + ncclGetErrorString(nr);
		return MC_ERR_CUDA;
	}
	for (int r = 0; r < c->n; r++) {
		const int rc = mc_exchange_sum(c->ctx[r], c->gathered[r], c->n);
		if (rc) {
			c->err = mc_last_error(c->ctx[r]);
			return rc;
		}
	}
	return MC_OK;
}
