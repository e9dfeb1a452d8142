"""Runs the reference's OWN numpy forward, /root/reference/src/python/compute_simple.py, unmodified, and saves the values of
its local variables: `python run_reference_python.py <reference script> <tokens_dir> <weights_dir> <out.npz>`.

The script only prints (truncated) arrays, so its main() is executed under sys.setprofile and the locals of that frame are
read when it returns: nothing of the reference is edited, copied or re-implemented here.  Used by
tests/test_ref_python_model.py and by make_ref_python_golden.py (which commits the outputs as a fixture)."""
import runpy
import sys

import numpy as np

WANT = ("K", "Q", "logits", "exp_approx", "attn", "O", "attn_out", "x_attn_res", "x_norm0", "ff_hidden", "x_norm1", "cls", "y_logit", "y_prob", "y_pred",
        "X_E", "X_F", "x_in")


def main():
    script, tokens_dir, weights_dir, out = sys.argv[1:5]
    captured = {}

    def hook(frame, event, arg):
        if event == "return" and frame.f_code.co_name == "main" and frame.f_code.co_filename == script:
            for name in WANT:
                if name in frame.f_locals:
                    captured[name] = np.asarray(frame.f_locals[name])

    sys.argv = [script, "--tokens_dir", tokens_dir, "--weights_dir", weights_dir]
    sys.setprofile(hook)
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        sys.setprofile(None)
    if "y_logit" not in captured:
        raise SystemExit("the reference script did not reach the end of main()")
    np.savez(out, **captured)


if __name__ == "__main__":
    main()
