import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from debug_perf import bench
import debug_tc
if __name__ == '__main__':
    debug_tc.run(0, 1, 32, 512, 32, bias=True, res_mode=1)
    debug_tc.run(0, 4, 64, 256, 64, bias=True)
    debug_tc.run(0, 1, 64, 512, 64)
    debug_tc.run(0, 1, 96, 512, 32, bias=True)
    debug_tc.run(3, 2, 128, 256, 256, bias=True)
    debug_tc.run(0, 3, 128, 200, 128, bias=True)      # partial tiles in both directions
    for shape in [(0, 1, 64, 256, 64), (0, 1, 32, 512, 32), (0, 1, 64, 512, 64), (0, 1, 128, 256, 128), (0, 4, 256, 128, 256), (0, 16, 256, 64, 256), (0, 16, 128, 128, 128)]:
        bench(*shape, mt2=1)
        bench(*shape, mt2=0)
