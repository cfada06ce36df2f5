#!/usr/bin/env python
"""Generate nmmo_b200/data/*_tasks.npz from the reference's own curriculum sources (run where /root/reference exists).

  heldout      neurips23_evaluation/heldout_evaluation_task.py:30-138 (63 specs) with the 2048-d fp16 embeddings of
               neurips23_evaluation/heldout_task_with_embedding.pkl, read with an engine-free Unpickler
  sample_eval  neurips23_evaluation/sample_eval_task_with_embedding.pkl (24 specs + embeddings)
  manual       curriculum_generation/manual_curriculum.py:53-314 (1609 specs; no embeddings: the reference's
               curriculum_with_embedding.pkl is listed in .MISSING_LARGE_BLOBS)

Every spec either becomes a row or is listed under `unsupported` with the reason.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from nmmo_b200.curriculum import DATA, load_spec_module, load_spec_pickle, spec_name, translate  # noqa: E402

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")


def save(name, specs, emb_from=None):
    rows, names, weights, emb, bad = translate(specs)
    if emb is None and emb_from is not None:
        by_name = {spec_name(s): s.embedding for s in emb_from}
        if all(n in by_name for n in names):
            emb = np.stack([np.asarray(by_name[n]).reshape(-1) for n in names]).astype(np.float16)
    out = {"rows": rows, "names": np.array(names), "weights": weights,
           "unsupported": np.array(bad if bad else np.zeros((0, 2), str))}
    if emb is not None:
        out["embed"] = emb
    DATA.mkdir(exist_ok=True)
    np.savez_compressed(DATA / f"{name}_tasks.npz", **out)
    print(f"{name}: {len(rows)} rows, {len(bad)} unsupported, embeddings {None if emb is None else emb.shape}")
    for n, why in bad[:8]:
        print("   unsupported:", n, "--", why)


heldout_pkl = load_spec_pickle(REF / "neurips23_evaluation" / "heldout_task_with_embedding.pkl")
heldout_py = load_spec_module(REF / "neurips23_evaluation" / "heldout_evaluation_task.py")
assert [spec_name(s) for s in heldout_py] == [spec_name(s) for s in heldout_pkl], "spec module and pickle disagree"
save("heldout", heldout_py, emb_from=heldout_pkl)
save("sample_eval", load_spec_pickle(REF / "neurips23_evaluation" / "sample_eval_task_with_embedding.pkl"))
save("manual", load_spec_module(REF / "curriculum_generation" / "manual_curriculum.py"))
