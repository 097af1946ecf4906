import sys, torch
sys.path.insert(0, '.')
from pyemir_b200 import abba
x = torch.randn((64, 2090, 3400), device='cuda')
labels = abba.labels_of(abba.build_full_set("ABBA", 1, 16))
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
sums = abba.partial_sums(x, labels)
print("partial_sums ms", timeit(lambda: abba.partial_sums(x, labels, sums=None)))
print("partial_sums (accumulate) ms", timeit(lambda: abba.partial_sums(x, labels, sums=sums)))
print("finish ms", timeit(lambda: abba.finish(sums, 32, 32)))
print("torch sum ms", timeit(lambda: x.sum(dim=0, dtype=torch.float32)))
print("torch sum f64 ms", timeit(lambda: x.sum(dim=0, dtype=torch.float64)))
