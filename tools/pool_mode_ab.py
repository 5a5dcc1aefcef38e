"""A/B of the two collapse kernels on a bench config: python tools/pool_mode_ab.py [cubes|shrec|humanseg|coseg]
mode 0 = by size (the default), 1 = always the round kernel, 2 = always the sequential kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import subprocess, json
name = sys.argv[1] if len(sys.argv) > 1 else 'cubes'
if len(sys.argv) > 2:
    from meshcnn_b200 import _lib
    _lib.lib().mcnn_debug_set_pool_mode(int(sys.argv[2]))
    import bench
    sys.argv = ['bench.py', '--config', name, '--no-e2e', '--no-cpu-baseline', '--steps', '60']
    bench.main()
else:
    for mode in (0, 1, 2):
        out = subprocess.run([sys.executable, __file__, name, str(mode)], capture_output=True, text=True).stdout.strip().splitlines()
        d = json.loads(out[-1])
        print('%s pool mode %d: %.1f meshes/s, %.3f ms/step' % (name, mode, d['value'], d['ms_per_step']))
