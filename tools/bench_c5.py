"""C5: 1024 utterances with N ~ U{64..512} phonemes (seed 7), all arriving at once, length-bucketed,
through the public batched call (host ids in, host waveforms out).  Prints one JSON line with the aggregate
audio-s/s and the p50 / p99 per-utterance completion latency.  One process per GPU under torchrun: the utterances are
LPT-partitioned over the ranks (shard.lpt_partition), each rank runs its own buckets, no collective on the data path.

  python tools/bench_c5.py [--utts 1024] [--batch 32]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c5.py

NOTE: written at the end of round 1 after the GPU budget was spent -- the scheduling helpers are unit-tested
(tests/test_host.py), the GPU loop below has not been run on hardware yet.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--quality", default="medium")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import sonata_b200
    from sonata_b200 import shard, voicegen, workload

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model = sonata_b200.from_config_path(voicegen.write_voice(voicegen.default_voice_dir(), args.quality), device=local)
    nph = workload.mixed_lengths(args.utts)
    ids = [workload.synthetic_ids(int(n), utt=u) for u, n in enumerate(nph)]
    mine = shard.lpt_partition([len(x) for x in ids], world)[rank] if world > 1 else list(range(len(ids)))
    my_ids = [ids[i] for i in mine]
    buckets = workload.length_buckets([len(x) for x in my_ids], args.batch)
    model.infer_batch_with_values([my_ids[i] for i in buckets[0]])            # warm-up (allocator, kernels)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    done, audio = [], []
    for b in buckets:
        auds = model.infer_batch_with_values([my_ids[i] for i in b])
        done.append(time.perf_counter() - t0)
        audio.append(sum(len(a) for a in auds) / 22050.0)
    p50, p99, agg = workload.completion_stats(buckets, done, audio)
    t = torch.tensor([done[-1], sum(audio), p50, p99], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        wall, total_audio, p50, p99 = float(mx[0]), float(sm[1]), float(mx[2]), float(mx[3])
    else:
        wall, total_audio = done[-1], sum(audio)
    if rank == 0:
        print(json.dumps({"workload": f"C5: {args.utts} utterances, N ~ U{{64..512}} phonemes, buckets of {args.batch}, "
                                      f"{world} GPU(s)", "metric": "audio-sec/sec", "value": total_audio / wall,
                          "unit": "audio-s/s", "n_gpus": world, "wall_s": wall, "audio_s": total_audio,
                          "latency_p50_s": p50, "latency_p99_s": p99,
                          "latency_note": "completion time of the utterance's bucket, all requests arriving at t=0; "
                                          "max over ranks of the per-rank percentiles when n_gpus > 1"}), flush=True)
    model.close()


if __name__ == "__main__":
    main()
