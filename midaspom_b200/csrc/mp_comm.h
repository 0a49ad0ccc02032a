// mp_comm.h -- NCCL calls of mp_comm.cu used by the sharded sweep in mp_engine.cu (all on the engine stream)
#pragma once
#include "mp_host.h"

int mp_comm_group_start(mp_engine *h);
int mp_comm_group_end(mp_engine *h);
int mp_comm_allgather(mp_engine *h, const void *send, void *recv, size_t bytes_per_rank);
int mp_comm_broadcast(mp_engine *h, void *buf, size_t bytes, int root);
