#!/usr/bin/env python3
"""BASELINE.json configs[4]: realignment microbench sweep — read length 100-250 bp x haplotype
length 300-1500 bp x 2-64 haplotypes per window (SURVEY.md §8d cfg 5: hap 0 random, haps 1..P-1
= hap 0 + one spiked variant, R = 512 reads sampled from the P haps, 5 % overhanging).
One JSON line per point: resident-batch pairs/s (CUDA events, L2 flushed), DP cells/s, chain
evaluations/s, and a bit-exact check of the first groups against the oracle.
usage: python tools/bench_sweep.py [--pairs 262144] [--steps 5] > profiles/r2_sweep.jsonl"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=262144)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--reads", type=int, default=512)
    args = ap.parse_args()
    import torch
    import oracle_lib as O
    from compare import compare_results
    from lancet2_b200 import abi, synth
    from lancet2_b200.realign import GpuRealigner
    gpu = GpuRealigner(0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for L in (100, 150, 250):
        for H in (300, 500, 1000, 1500):
            for P in (2, 4, 8, 16, 32, 64):
                n_groups = max(2, args.pairs // (args.reads * P))
                groups = synth.make_groups(1000 * L + H + P, n_groups, read_len=L, hap_len=H, n_haps=P, n_reads=args.reads)
                batch = abi.Batch(groups)
                chk = abi.Batch(groups[:2])
                want, _ = O.oracle_genotype(chk, gpu.params, n_threads=os.cpu_count() or 1)
                got, _ = gpu.genotype_batch(chk)
                ok = not compare_results(chk, want, got)
                gpu.upload_packed(abi.PackedBatch(groups, gpu.lib))  # the product path: packed wire format, unpack inside the step
                for _ in range(3):
                    gpu.run_resident()
                ms, st = 0.0, None
                for _ in range(args.steps):
                    flush.fill_(1)
                    torch.cuda.synchronize()
                    st = gpu.run_resident()
                    ms += st.ms_kernels
                sec = ms * 1e-3 / args.steps
                print(json.dumps({"read_len": L, "hap_len": H, "haps_per_group": P, "groups": n_groups, "pairs": batch.n_pairs,
                                  "pairs_per_s": batch.n_pairs / sec, "ms_per_step": sec * 1e3,
                                  "aligned_frac": st.n_aligned / max(1, st.n_pairs),
                                  "gcups_computed": st.dp_cells / sec / 1e9, "gcups_reference_rectangles": st.dp_cells_full / sec / 1e9,
                                  "chain_gevals_per_s": st.chain_evals / sec / 1e9,
                                  "oracle_checked_pairs": chk.n_pairs, "bit_exact": ok}), flush=True)
    gpu.close()


if __name__ == "__main__":
    main()
