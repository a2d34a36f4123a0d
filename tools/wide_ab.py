"""A/B matrix for the wide-key count (k >= 48) on the headline read set: every configuration of the planner / kernel switches
runs `count` at each k on the same device-resident reads, its edge records are compared bit for bit with the first
configuration's (torch on the device: no 1.6 GB downloads), and the stage times are written to gpurun_out/<tag>.json.

  python tools/wide_ab.py --klist 119,141 --pairs 16666667 --configs old,default --tag r2ab_wide_ab
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OLD = {"MFSDBG_READS_COMPACT": "0", "MFSDBG_L1_CAP_WC": "8", "MFSDBG_L1_SEGK_WC": "4000", "MFSDBG_L2_NBCAP_W": "1024", "MFSDBG_CW_CHUNK": "0",
       "MFSDBG_READS_COMPACT_V": "1"}
CONFIGS = {
    # name: env.  "old" = the build before the compacted scatter (gpurun_out/r2aa_wide_ab.json measured the steps in between)
    "old": OLD,
    "default": {},
    "walk": {"MFSDBG_READS_COMPACT": "0"},
    "compact": {"MFSDBG_READS_COMPACT": "2"},
    "compact_l1_8": {"MFSDBG_READS_COMPACT": "2", "MFSDBG_L1_CAP_WC": "8", "MFSDBG_L1_SEGK_WC": "4000"},
    "nb1024": {"MFSDBG_L2_NBCAP_W": "1024"},
    "cw0": {"MFSDBG_CW_CHUNK": "0"},
    "cw1": {"MFSDBG_CW_CHUNK": "1"},
    "v1": {"MFSDBG_READS_COMPACT_V": "1"},
    "v2": {"MFSDBG_READS_COMPACT_V": "2"},
    "v2_cw1": {"MFSDBG_READS_COMPACT_V": "2", "MFSDBG_CW_CHUNK": "1"},
}
KNOBS = sorted(OLD)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--klist", default="59,79,99,119,141")
    ap.add_argument("--pairs", type=int, default=16_666_667)
    ap.add_argument("--min-count", type=int, default=2)
    ap.add_argument("--configs", default=",".join(CONFIGS))
    ap.add_argument("--tag", default="wide_ab")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import torch
    from mitoflex_b200 import lib
    ctx = lib.Context(0)
    ctx.set_profiling(True)
    reads = ctx.synth(n_pairs=args.pairs, seed=1002)
    out = {"pairs": args.pairs, "bases": reads.n_bases, "min_count": args.min_count, "runs": []}

    def dump():   # after every run: a crash later on keeps what was measured
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", args.tag + ".json"), "w") as f:
            json.dump(out, f, indent=1)

    class _DevArray:   # the library's edge buffer seen by torch (cloned at once: the next call reuses the memory)
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (int(ptr), False), "version": 2}

    def edges_tensor(e):
        n = e.s.n_edges * e.s.words_per_edge
        if n == 0:
            return torch.empty(0, dtype=torch.int32, device="cuda:0")
        return torch.as_tensor(_DevArray(e.s.edges, n), device="cuda:0").clone()

    for k in [int(x) for x in args.klist.split(",")]:
        ref = None
        for name in args.configs.split(","):
            for kn in KNOBS:
                os.environ.pop(kn, None)
            os.environ.update(CONFIGS[name])
            rec = {"k": k, "config": name}
            try:
                best = None
                for _ in range(args.reps):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    e = ctx.count(reads, k, args.min_count)
                    torch.cuda.synchronize()
                    ms = 1e3 * (time.perf_counter() - t0)
                    if best is None or ms < best:
                        best = ms
                        prof = ctx.last_profile()
                rec["count_ms"] = round(best, 2)
                rec["n_edges"] = int(e.n)
                rec["stages_ms"] = {k2: round(v, 2) for k2, v in sorted(prof.items(), key=lambda kv: -kv[1])[:8]}
                t = edges_tensor(e)
                if ref is None:
                    ref = t
                    rec["equal_to_first"] = True
                else:
                    rec["equal_to_first"] = bool(t.shape == ref.shape and torch.equal(t, ref))
                    del t
            except Exception as ex:   # a configuration the planner rejects must not end the matrix
                rec["error"] = repr(ex)[:300]
            print(json.dumps(rec), flush=True)
            out["runs"].append(rec)
            dump()
        del ref
        torch.cuda.empty_cache()
    bad = [r for r in out["runs"] if r.get("equal_to_first") is False or "error" in r]
    print("MISMATCH/ERROR:" if bad else "ALL EQUAL", json.dumps(bad)[:2000])


if __name__ == "__main__":
    main()
