"""Run under torch.distributed.run: mode B (one merged batch) through `sharded.merge_distributed`.

Backend is taken from PA_DIST_BACKEND: "gloo" (two ranks sharing one GPU, payload staged through host memory — the
orchestration test the driver can run on a 1-GPU box) or "nccl" (one GPU per rank, NVLink send/recv). Every rank builds the
same full workload, keeps its pid-hash shard, hashes/deduplicates it on its GPU; rank 0 merges and compares the result with
the CPU oracle on the unsharded stream.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py  # noqa: E402
from parca_agent_b200 import abi, lib, sharded, synth  # noqa: E402


def main():
    backend = os.environ.get("PA_DIST_BACKEND", "gloo")
    local = int(os.environ.get("LOCAL_RANK", "0")) if backend == "nccl" else 0
    torch.cuda.set_device(local)
    dist.init_process_group(backend)
    rank, world = dist.get_rank(), dist.get_world_size()
    for schema in (abi.PA_SCHEMA_V2, abi.PA_SCHEMA_V1):
        for w in (synth.edge_workload(seed=33, n=5000, hash_mode=abi.PA_HASH_PROVIDED), synth.config3(n=80_000, u=5_000, p=4_096, npids=96, lsets=6)):
            idx = sharded.shard_rows(w, world)[rank]
            part = w.rows(idx)
            part.schema = abi.PA_SCHEMA_V2
            a = lib.from_workload(part, device=local)
            lib.load(a, part)
            a.stage()
            a.process()
            merged = None
            if rank == 0:
                merged = lib.Aggregator(device=local, hash_mode=abi.PA_HASH_PROVIDED, label_flags=w.label_flags, samples_per_second=w.samples_per_second,
                                        external_labels=w.external_labels, max_samples=w.n, max_frames=max(w.n_frame_ids, 1), schema=schema)
                merged.register_strings(w.strings[1:])
                merged.register_frames(w.frames)
                merged.register_labelsets(w.labelsets)
            res = sharded.merge_distributed(a, idx, merged, dst=0, device=local)
            if rank == 0:
                w.schema = schema
                want, st = oracle_py.run(w)
                assert res.ipc_bytes() == want, "merged batch differs from the unsharded oracle (%s, schema %d)" % (w.name, schema)
                assert res.n_rows == w.n and res.n_unique_stacks == st["unique_stacks"]
                merged.close()
            a.close()
            dist.barrier()
    if rank == 0:
        print("merge-check ok world=%d backend=%s" % (world, backend))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
