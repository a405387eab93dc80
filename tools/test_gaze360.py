"""Evaluation entry point of the B200 backend: same arguments as the reference's tools/test_gaze360_gaze.py
(config, checkpoint, --json, --root, --device, --cfg-options), same results JSON (results/results_<config>_<json>),
but MANY clips per forward, the image pipeline on the GPU, and -- under torchrun -- clips sharded over the ranks with
one all-gather (what the reference's dist_test.sh path would do if its test-mode dataset existed).

    python tools/test_gaze360.py configs/multiclue_gaze/multiclue_gaze_r50_gaze360.py ckpt.pth
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/test_gaze360.py <config> <ckpt>
"""
import json
import os
import sys
from argparse import ArgumentParser

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcgaze_b200 import evaluate as ev  # noqa: E402
from mcgaze_b200.apis import init_detector  # noqa: E402
from mcgaze_b200.compat import DictAction  # noqa: E402
from mcgaze_b200.pipeline import GpuTestPipeline  # noqa: E402


def main():
    ap = ArgumentParser()
    ap.add_argument('config')
    ap.add_argument('checkpoint')
    ap.add_argument('--json', default='data/gaze360/test.json')
    ap.add_argument('--root', default='data/gaze360/test_rawframes/')
    ap.add_argument('--device', default=None)
    ap.add_argument('--cfg-options', nargs='+', action=DictAction)
    ap.add_argument('--clips-per-batch', type=int, default=32)
    ap.add_argument('--workers', type=int, default=8, help='decode threads (frames of the next batch load while the GPU runs)')
    ap.add_argument('--seed', type=int, default=None, help='pins the CenterCrop draws (the reference run is unseeded)')
    ap.add_argument('--scorer', default=None, choices=['gaze360', 'l2cs'],
                    help='which of the reference scorers to print (default: l2cs when the config name says so)')
    ap.add_argument('--shard', default='video', choices=['video', 'clip'],
                    help='multi-GPU sharding: videos (merge + MAE on each device, one all-reduce) or clips (one all-gather)')
    ap.add_argument('--decode', default='gpu', choices=['gpu', 'host'],
                    help='PNG frames decoded on the device (mcg_png_decode; the loader threads only read files) or by cv2 '
                         'on the host like LoadImageFromFile; other formats always take the host path')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    device = args.device or f'cuda:{local}'
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device(device))
    model = init_detector(args.config, args.checkpoint, device=device, cfg_options=args.cfg_options)
    pipe = GpuTestPipeline(model.cfg.data.test.pipeline, device=int(device.split(':')[1]), seed=args.seed)
    ds = ev.Gaze360ClipDataset(args.json, img_prefix=args.root, decode=args.decode)
    scorer = args.scorer or ('l2cs' if 'l2cs' in os.path.basename(args.config) else 'gaze360')
    sharded = None
    if args.shard == 'video':
        # videos sharded over the ranks (all of them on the one rank of a single-GPU run): overlap merge and MAE of the
        # fused gaze on each device, one all-reduce of 6 doubles, one all-gather of the merged rows for the JSON
        sharded = ev.multi_gpu_test_videos(model, ds, pipe, args.clips_per_batch, workers=args.workers, variant=scorer)
        records = ev.records_from_merged(ds, sharded['merged'])
    elif world > 1:
        rows = ev.multi_gpu_test(model, ds, pipe, args.clips_per_batch, device=device, workers=args.workers)
        records, _ = ev.videos_from_clips(ds, rows)
    else:
        rows = ev.single_gpu_test(model, ds, pipe, args.clips_per_batch, workers=args.workers)
        records, _ = ev.videos_from_clips(ds, rows)
    if int(os.environ.get('RANK', '0')) == 0:
        os.makedirs('results', exist_ok=True)
        out = os.path.join('results', f'results_{os.path.basename(args.config)[:-3]}_{os.path.basename(args.json)}')
        json.dump(records, open(out, 'w'))
        print('wrote', out)
        if ds.anno.get('annotations'):
            if sharded is not None and 'mae' in sharded:
                m = sharded['mae']
                print(f"device scorer ({scorer}), fusion_gazes: MAE 360 {m['mae_360']:.2f}  front-180 {m['mae_front90']:.2f}  "
                      f"front-20 {m['mae_front20']:.2f}")
            for key in ('fusion_gazes', 'face_gazes', 'eyes_gazes', 'head_gazes'):
                m = ev.evaluate(ds, records, key, variant=scorer)
                print(f"{key}: MAE 360 {m['mae_360']:.2f}  front-180 {m['mae_front90']:.2f}  front-20 {m['mae_front20']:.2f}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
