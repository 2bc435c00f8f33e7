/*
 * mc_comm.h -- NCCL exchange step of an individual-sharded fit driven from ONE
 * host process (the C command line with --gpus N).  Lives in its own library
 * (libmc_comm.so, linked against the system NCCL) so that programs which bring
 * their own collective layer -- bench.py uses torch.distributed -- do not load a
 * second NCCL.
 *
 * The reference has no counterpart: it is single-process, single-thread
 * (SURVEY.md section 5).  What is exchanged per EM step is the buffer of
 * mc_exchange_buffer(): [K*T allele-count sums | ll | K pooled-eta sums].
 */
#ifndef MC_COMM_H
#define MC_COMM_H

#include "mc_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mc_comm mc_comm;

/* one communicator over the contexts of one process; ctxs[r] must live on
 * distinct devices and hold models of identical K and T */
int mc_comm_create(mc_comm **comm, mc_ctx **ctxs, int n);
void mc_comm_destroy(mc_comm *comm);
/* Sum of every context's exchange buffer over NVLink as a deterministic
 * reduce-scatter + all-gather: grouped ncclSend / ncclRecv move slice j of
 * every buffer to device j, mc_exchange_sum_slice adds the copies in rank
 * order, one in-place ncclAllGather returns the totals -- afterwards all
 * contexts hold bit-identical totals, and the bytes per device do not grow
 * with the number of devices.  Every call runs on its context's stream;
 * asynchronous with respect to the host. */
int mc_comm_exchange(mc_comm *comm);
const char *mc_comm_last_error(const mc_comm *comm);

#ifdef __cplusplus
}
#endif

#endif
