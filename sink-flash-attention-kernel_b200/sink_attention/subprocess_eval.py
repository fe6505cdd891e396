"""``subprocess_generate``: run ``model.generate()`` in a fresh Python process.

Process plumbing only (no arithmetic) -- kept so the reference's 12-name API surface
(sink_attention/__init__.py:1-28, subprocess_eval.py) stays importable.  The child loads the model,
patches it with ``patch_for_generation`` and writes the generated ids to a JSON file; on a non-zero
exit the call is retried with more visible GPUs (1 -> 2 -> 4), as the reference does.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile
from typing import List, Optional, Sequence

_CHILD = r"""
import json, sys, torch
cfg = json.load(open(sys.argv[1]))
sys.path[:0] = cfg["sys_path"]
from transformers import AutoModelForCausalLM, AutoTokenizer
from sink_attention import patch_for_generation
tok = AutoTokenizer.from_pretrained(cfg["model"])
model = AutoModelForCausalLM.from_pretrained(cfg["model"], torch_dtype=getattr(torch, cfg["dtype"]),
                                             device_map="auto", attn_implementation="flash_attention_2")
outs = []
for prompt in cfg["prompts"]:
    cache = patch_for_generation(model, num_sink=cfg["num_sink"], window_size=cfg["window_size"])
    ids = tok(prompt, return_tensors="pt").input_ids.to(model.device)
    gen = model.generate(ids, past_key_values=cache, max_new_tokens=cfg["max_new_tokens"], do_sample=False)
    outs.append(tok.decode(gen[0, ids.shape[1]:], skip_special_tokens=True))
json.dump({"outputs": outs}, open(cfg["out"], "w"))
"""


def subprocess_generate(model_name_or_path: str, prompts: Sequence[str], max_new_tokens: int = 128,
                        num_sink: int = 4, window_size: int = 4096, dtype: str = "bfloat16",
                        gpu_counts: Sequence[int] = (1, 2, 4), timeout: Optional[float] = None) -> List[str]:
    """Generate completions for ``prompts`` in a child process; returns the decoded strings."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    last_err = ""
    for n_gpu in gpu_counts:
        with tempfile.TemporaryDirectory() as tmp:
            cfg = {"model": model_name_or_path, "prompts": list(prompts), "max_new_tokens": max_new_tokens,
                   "num_sink": num_sink, "window_size": window_size, "dtype": dtype,
                   "out": os.path.join(tmp, "out.json"), "sys_path": [here]}
            cfg_path = os.path.join(tmp, "cfg.json")
            with open(cfg_path, "w") as f:
                json.dump(cfg, f)
            env = dict(os.environ)
            env["CUDA_VISIBLE_DEVICES"] = ",".join(str(i) for i in range(n_gpu))
            proc = subprocess.run([sys.executable, "-c", _CHILD, cfg_path], env=env, capture_output=True,
                                  text=True, timeout=timeout)
            if proc.returncode == 0 and os.path.exists(cfg["out"]):
                with open(cfg["out"]) as f:
                    return json.load(f)["outputs"]
            last_err = proc.stderr[-2000:]
    raise RuntimeError(f"subprocess_generate failed on GPU counts {tuple(gpu_counts)}:\n{last_err}")
