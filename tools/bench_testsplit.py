"""BASELINE configs[3] stand-in: the Gaze360 test split (517 videos, the REAL video-length list of the shipped results
JSON -> 6365 clips, 6237 of 7 frames) with synthetic pixels, through the whole evaluation driver: clip slicing, frame
loading (synthetic: a cached random uint8 frame per call, i.e. PNG decoding excluded), GPU pipeline, batched forward,
result read-back, (under torchrun) sharding + one all-gather, overlap merge, JSON records.

    python tools/bench_testsplit.py [clips_per_batch] [workers] [src_h src_w]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_testsplit.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mcgaze_b200 import evaluate as ev  # noqa: E402
from mcgaze_b200.apis import init_detector  # noqa: E402
from mcgaze_b200.pipeline import GpuTestPipeline  # noqa: E402
from oracle import mcgaze_oracle as O  # noqa: E402  (seeded synthetic checkpoint only)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    mode = 'video'
    argv = sys.argv[1:]
    if argv and argv[0] in ('video', 'clip'):
        mode, argv = argv[0], argv[1:]
    a = [int(v) for v in argv]
    cpb = a[0] if len(a) > 0 else 32
    workers = a[1] if len(a) > 1 else 8
    sh, sw = (a[2], a[3]) if len(a) > 3 else (300, 300)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    device = f'cuda:{local}'
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device(device))
    lengths = np.load(os.path.join(ROOT, 'tests/golden/golden_gaze360_results.npz'))['lengths'].tolist()
    grng = np.random.default_rng(1)
    anno = dict(videos=[dict(id=i + 1, file_names=[f'{i:04d}/{t:05d}.png' for t in range(L)]) for i, L in enumerate(lengths)],
                annotations=[dict(gaze=grng.normal(size=(L, 3)).tolist()) for L in lengths])
    rng = np.random.default_rng(0)
    bank = [rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8) for _ in range(64)]
    ds = ev.Gaze360ClipDataset(anno, loader=lambda p: bank[hash(p) % 64])
    model = init_detector(os.path.join(ROOT, 'configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py'), None, device)
    model.load_state_dict(O.make_state_dict(0))
    pipe = GpuTestPipeline(model.cfg.data.test.pipeline, device=local, seed=0)
    ev.run_clips(model, ds, pipe, list(range(64)), cpb, workers)            # warm-up: engine, plans
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    mae = None
    if mode == 'video':      # videos sharded, merge + MAE on each device, one all-reduce (+ one all-gather for the JSON)
        out = ev.multi_gpu_test_videos(model, ds, pipe, cpb, workers=workers)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        records = ev.records_from_merged(ds, out['merged'])
        mae = out['mae']
    else:                    # clips sharded, one all-gather of the per-clip rows, merge on the host
        rows = ev.multi_gpu_test(model, ds, pipe, cpb, device=device, workers=workers) if world > 1 else \
            ev.single_gpu_test(model, ds, pipe, cpb, workers)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        records, _ = ev.videos_from_clips(ds, rows)
    t2 = time.perf_counter()
    if rank == 0:
        print(json.dumps(dict(workload='Gaze360 test split stand-in (BASELINE configs[3])', videos=len(lengths), frames=int(sum(lengths)),
                              clips=len(ds), n_gpus=world, shard=mode, mae=mae, clips_per_batch=cpb, workers=workers, source=[sh, sw],
                              forward_s=t1 - t0, merge_json_s=t2 - t1, clips_per_s=len(ds) / (t1 - t0),
                              note='wall clock incl. host frame hand-over, H2D of uint8, GPU pipeline, forward, D2H, gather; '
                                   'PNG decoding excluded (frames come from a RAM bank); short clips (T < 7) run in their own batches')))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
