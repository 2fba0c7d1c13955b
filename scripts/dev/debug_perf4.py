import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from debug_perf import bench
if __name__ == '__main__':
    for shape in [(0, 1, 32, 512, 32), (0, 1, 64, 512, 64), (0, 1, 64, 256, 64), (0, 1, 128, 256, 128), (0, 1, 128, 128, 128), (0, 1, 256, 64, 256),
                  (0, 1, 256, 32, 256), (0, 1, 512, 16, 512), (0, 1, 256, 16, 256), (3, 1, 256, 16, 768), (0, 1, 96, 512, 32), (0, 16, 256, 64, 256), (0, 16, 64, 256, 64)]:
        bench(*shape, half=0)
        bench(*shape, half=1)
