import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from debug_perf import bench
if __name__ == '__main__':
    print('THA4_TC_STAGES =', os.environ.get('THA4_TC_STAGES'))
    bench(3, 1, 32, 16, 32, ksplit=1, cluster=0)       # minimal: 1 k-block, 2 CTAs
    bench(3, 1, 256, 16, 256, ksplit=1, cluster=0)     # 8 k-blocks, 2 CTAs
    bench(0, 1, 256, 16, 256, ksplit=8, cluster=0)
    bench(0, 1, 256, 16, 256, ksplit=4, cluster=1)
    bench(0, 1, 256, 16, 256, ksplit=8, cluster=1)
    bench(0, 1, 512, 16, 512, ksplit=16, cluster=0)
    bench(0, 1, 128, 128, 128, ksplit=1, cluster=0)
