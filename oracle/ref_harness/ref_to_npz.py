"""inputs + reference outputs (.b2ref) -> tests/golden/ref_<tag>.npz.   python ref_to_npz.py <inputs dir> <outputs dir> <golden dir> [source note]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import b2ref  # noqa: E402


def convert(in_dir, out_dir, golden_dir, source="rmcl @ reference checkout + rmagine + Embree (oracle/ref_harness/gen_ref_golden)"):
    os.makedirs(golden_dir, exist_ok=True)
    done = []
    for tag in ("c1", "pin"):
        pi, po = os.path.join(in_dir, tag + ".b2ref"), os.path.join(out_dir, tag + ".b2ref")
        if not (os.path.exists(pi) and os.path.exists(po)):
            continue
        rec = {("in." + k): v for k, v in b2ref.read(pi).items()}
        rec.update({("ref." + k): v for k, v in b2ref.read(po).items()})
        rec["source"] = np.array(source)
        np.savez_compressed(os.path.join(golden_dir, f"ref_{tag}.npz"), **rec)
        done.append(tag)
    return done


if __name__ == "__main__":
    print("written:", convert(sys.argv[1], sys.argv[2], sys.argv[3], *(sys.argv[4:5])))
