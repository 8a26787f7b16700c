"""One-shot GPU bring-up check (not a test): every seeded case, ours vs oracle vs reference kernels,
without stopping at the first mismatch.  Writes gpurun_out/gpu_check.json."""
import json
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import cases, runners  # noqa: E402


def main():
    import torch
    from bdm_b200 import backend as ours
    from oracle import build_ref
    ref = build_ref.load_ref()
    print("device", torch.cuda.get_device_name(0), "ref ext:", ref is not None, flush=True)
    res = {}
    for name in cases.GOLDEN_CASES:
        inp = cases.build_case(name)
        entry = {}
        try:
            got = runners.run_backend(name, ours, inp)
            torch.cuda.synchronize()
        except Exception:
            entry["ours"] = "EXC " + traceback.format_exc()[-400:]
            res[name] = entry
            print(name, entry, flush=True)
            continue
        for tag, fn in (("oracle", lambda: runners.run_oracle(name, inp)),
                        ("ref", (lambda: runners.run_backend(name, ref, inp)) if ref is not None else None)):
            if fn is None:
                continue
            try:
                want = fn()
                runners.compare(name, got, want)
                entry[tag] = "ok"
                if tag == "ref":  # also record whether floats are bit-identical to the reference kernels
                    entry["bitexact_float_keys"] = {k: bool((got[k] == want[k]).all()) for k in want
                                                    if want[k].dtype.kind == "f"}
            except AssertionError as e:
                entry[tag] = "FAIL " + str(e)[:300]
            except Exception:
                entry[tag] = "EXC " + traceback.format_exc()[-400:]
        # oracle vs reference kernels (pins the oracle)
        if ref is not None:
            try:
                runners.compare(name, runners.run_oracle(name, inp), runners.run_backend(name, ref, inp))
                entry["oracle_vs_ref"] = "ok"
            except AssertionError as e:
                entry["oracle_vs_ref"] = "FAIL " + str(e)[:300]
        res[name] = entry
        print(name, entry, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
