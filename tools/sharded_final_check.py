#!/usr/bin/env python
"""torchrun check of the date-ordered weighting across GPUs (development tool):
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_final_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from document_search_engine_b200 import DescDateBM25F
from document_search_engine_b200.corpus import config_corpus
from document_search_engine_b200.distributed import ShardedSearcher
from oracle.numpy_oracle import NumpyOracle
from tests.parity import assert_query_parity
from tests.test_date_final import FINAL_TOL, date_queries, dated_corpus

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
torch.cuda.set_stream(torch.cuda.Stream())
ix = dated_corpus(config_corpus(1, device="cpu"))
qs = date_queries(200, 3)
ss = ShardedSearcher(ix, rank=rank, world=world, weighting=DescDateBM25F)
batch = ss.pack(qs)
for k in (10, 100):
    final, docids, counts, totals = ss.search_packed(batch, k)
    if rank == 0:
        o = NumpyOracle(ix, final_add=DescDateBM25F().doc_final_terms(ix))
        for i, q in enumerate(qs):
            n = int(counts[i])
            assert_query_parity(o, q, list(zip(final[i, :n].tolist(), docids[i, :n].tolist())), int(totals[i]), k,
                                ctx="query %d" % i, abs_tol=FINAL_TOL)
        print("sharded final k=%d: %d queries match the oracle on %d GPUs" % (k, len(qs), world))
dist.barrier()
dist.destroy_process_group()
