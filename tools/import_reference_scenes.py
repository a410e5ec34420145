#!/usr/bin/env python
"""Import the reference's scene DATA (not code) into this repo.

Reads the ten scene definition files `plb/envs/*.yml` and the fifty 64^3
target-density grids `plb/envs/assets/*.npy` from a PlasticineLab checkout and
writes

  plasticinelab_b200/envs/scenes.json          the parsed YAML trees, verbatim values
  plasticinelab_b200/envs/assets/targets.npz   sparse (flat index, float64 value) per grid

so that the engine, the tests and bench.py work on a box where the checkout is
absent (the GPU box).  The grids are stored losslessly (float64 values of the
non-zero voxels); `plasticinelab_b200.envs.assets.load_target` rebuilds the
dense array bit-exactly.

Usage:  python tools/import_reference_scenes.py [/root/reference]
"""
import json
import os
import sys

import numpy as np
import yaml

SCENES = ['move', 'torus', 'rope', 'writer', 'pinch', 'rollingpin', 'chopsticks', 'table',
          'triplemove', 'assembly']


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    here = os.path.dirname(os.path.abspath(__file__))
    out_dir = os.path.join(here, '..', 'plasticinelab_b200', 'envs')
    scenes = {}
    for name in SCENES:
        with open(os.path.join(ref, 'plb', 'envs', name + '.yml')) as f:
            scenes[name] = yaml.safe_load(f)
    with open(os.path.join(out_dir, 'scenes.json'), 'w') as f:
        json.dump(scenes, f, indent=1, sort_keys=True)

    asset_dir = os.path.join(ref, 'plb', 'envs', 'assets')
    packed = {}
    for fn in sorted(os.listdir(asset_dir)):
        if not fn.endswith('.npy'):
            continue
        a = np.load(os.path.join(asset_dir, fn))
        assert a.dtype == np.float64 and a.shape == (64, 64, 64), (fn, a.dtype, a.shape)
        flat = a.reshape(-1)
        idx = np.nonzero(flat)[0].astype(np.int32)
        key = fn[:-4]
        packed[key + '.idx'] = idx
        packed[key + '.val'] = flat[idx]
    os.makedirs(os.path.join(out_dir, 'assets'), exist_ok=True)
    np.savez_compressed(os.path.join(out_dir, 'assets', 'targets.npz'), **packed)
    print('scenes:', len(scenes), 'targets:', len(packed) // 2)


if __name__ == '__main__':
    main()
