"""create a synthetic workflow directory (frames + depth maps + config.json) for driver runs"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'video-stereo-converter_b200')]
import cv2, numpy as np
from vsc_b200.synthetic import make_pair
wf, n, h, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
for d in ('frames', 'depth_maps', 'sbs'):
    os.makedirs(os.path.join(wf, d), exist_ok=True)
cfg = {'input_video': 'in.mkv', 'output_video': 'out.mkv',
       'directories': {'frames': 'frames', 'depth_maps': 'depth_maps', 'sbs': 'sbs', 'chunks': 'chunks'},
       'stereo': {'max_disparity': 50.0, 'convergence': -10.0, 'super_sampling': 3.0, 'edge_softness': 20.0,
                  'artifact_smoothing': 1.0, 'depth_gamma': 0.2, 'sharpen': 14.0},
       'depth': {'save_16bit': True}, 'encoding': {'crf': 19, 'preset': 'slow'},
       'free_space': {'sbs_generator': 'none', 'chunk_generator': 'none'}}
json.dump(cfg, open(os.path.join(wf, 'config.json'), 'w'))
distinct = int(sys.argv[5]) if len(sys.argv) > 5 else n      # frames beyond `distinct` are symlinks to the first ones
for i in range(n):
    fp, dp = os.path.join(wf, 'frames', f'frame_{i:06d}.png'), os.path.join(wf, 'depth_maps', f'depth_frame_{i:06d}.tif')
    if i >= distinct:
        os.symlink(os.path.join(wf, 'frames', f'frame_{i % distinct:06d}.png'), fp)
        os.symlink(os.path.join(wf, 'depth_maps', f'depth_frame_{i % distinct:06d}.tif'), dp)
        continue
    rgb, depth = make_pair(h, w, seed=i % 8, depth_dtype=np.uint16)
    cv2.imwrite(fp, cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR))
    cv2.imwrite(dp, depth)
print('workflow', wf, n, 'frames', w, 'x', h)
