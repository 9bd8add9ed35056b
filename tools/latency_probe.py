import sys, time
sys.path.insert(0, '.')
from lancet2_b200 import abi, synth
from lancet2_b200.realign import GpuRealigner
gs = synth.make_region_groups(42, ref_len=100_000)
gpu = GpuRealigner(0)
for n in (1, 4, 16):
    b = abi.Batch(gs[:n]); res = abi.Result(b, 1 << 16)
    for _ in range(20): gpu.genotype_batch(b, result=res, want_aln=False)
    t0 = time.perf_counter()
    for _ in range(200): _, st = gpu.genotype_batch(b, result=res, want_aln=False)
    dt = (time.perf_counter() - t0) / 200
    print(n, b.n_pairs, "wall_ms", round(dt * 1e3, 3), "h2d", round(st.ms_h2d, 3), "kern", round(st.ms_kernels, 3), "d2h", round(st.ms_d2h, 3),
          {k: round(getattr(st, k), 3) for k in ("ms_k_index", "ms_k_sketch", "ms_k_map", "ms_k_ext", "ms_k_assign")})
