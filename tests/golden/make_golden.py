#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the CNN hot path into small fixtures.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Outputs (committed): tests/golden/conv2d.npz, conv_bias.npz, literals.json, ref_symbols.json

Sources (all under /root/reference):
  test_data_json/conv2d.json, conv_bias.json                  (PyTorch-generated, serde Matrix form)
  zenu-matrix/src/nn/batch_norm.rs:787-1077                    BN fwd-train / bwd / inference literals
  zenu-autograd/src/nn/batch_norm.rs:313-737                   BN fwd+bwd 2x3x4x4 via autograd, momentum 0.1
  zenu-matrix/src/operation/mul.rs:307-337                     GEMM 3x4 . 4x5
  zenu-matrix/src/operation/relu.rs:255-297                    ReLU / mask
  zenu-optimizer/tests/net_test.rs:82-260                      SGD / Adam / AdamW one-step literals
  zenu-cuda/src/cudnn/graph_conv.rs:433-623                    conv fwd 1x2x4x4 * 3x2x3x3 literals
  zenu-cuda-kernel-sys/kernel/*.h, zenu-cudnn-frontend-wrapper-sys/.../cudnn_frontend_wrapper.h
                                                               every extern "C" prototype of the two FFI surfaces (ref_symbols.json)
"""
import json
import os
import re

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

NUM = r"-?\d+\.?\d*(?:[eE][-+]?\d+)?"


def lines(path, lo, hi):
    with open(os.path.join(REF, path)) as f:
        return "".join(f.readlines()[lo - 1:hi])


def named_arrays(src):
    """Find `let NAME = vec![ ... ];` or `let NAME = [ ... ];` blocks, return {name: [floats]} in order."""
    out = {}
    for m in re.finditer(r"let\s+(?:mut\s+)?(\w+)\s*(?::[^=]+)?=\s*(?:vec!)?\[([^\]]*)\]", src, re.S):
        name, body = m.group(1), m.group(2)
        vals = [float(x) for x in re.findall(NUM, body)]
        if vals:
            out.setdefault(name, vals)
    return out


def serde_matrix(d):
    a = np.asarray(d["data"], dtype=np.float32)
    return a.reshape(d["shape"])


def main():
    for name in ("conv2d", "conv_bias"):
        with open(os.path.join(REF, "test_data_json", name + ".json")) as f:
            d = json.load(f)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **{k: serde_matrix(v) for k, v in d.items()})

    lit = {}
    # --- BN matrix-level: forward train (zenu-matrix/src/nn/batch_norm.rs:787-850)
    a = named_arrays(lines("zenu-matrix/src/nn/batch_norm.rs", 787, 850))
    x16 = re.findall(NUM, lines("zenu-matrix/src/nn/batch_norm.rs", 788, 806))
    lit["bn_fwd_train"] = {
        "shape": [2, 2, 2, 2], "momentum": 0.1,
        "x": [float(v) for v in x16 if "." in v][:16],
        "y": a["y"], "running_mean": a["running_mean"], "running_variance": a["running_variance"],
        "saved_mean": a["saved_mean"], "saved_inv_std": a["saved_variance"],
        "scale": a["scale"], "bias": a["bias"], "tol": 2e-4,
    }
    # --- BN backward (:895-993)
    a = named_arrays(lines("zenu-matrix/src/nn/batch_norm.rs", 895, 993))
    lit["bn_bwd"] = {
        "shape": [2, 2, 2, 2], "x": a["x"], "y_grad": a["y_grad"], "saved_mean": a["saved_mean"],
        "saved_inv_std": a["saved_variance"], "scale": a["scale"], "x_grad": a["x_grad_ans"],
        "scale_grad": a["scale_grad_ans"], "bias_grad": a["bias_grad_ans"], "tol": 2e-4,
    }
    # --- BN inference (:1027-1077)
    a = named_arrays(lines("zenu-matrix/src/nn/batch_norm.rs", 1027, 1077))
    lit["bn_infer"] = {"shape": [2, 2, 2, 2], "x": a["x"], "y": a["y"], "mean": a["mean"],
                       "variance": a["variance"], "scale": a["scale"], "bias": a["bias"], "tol": 3e-3}
    # --- BN via autograd 2x3x4x4 (zenu-autograd/src/nn/batch_norm.rs:313-737)
    a = named_arrays(lines("zenu-autograd/src/nn/batch_norm.rs", 313, 737))
    lit["bn_autograd"] = {
        "shape": [2, 3, 4, 4], "momentum": 0.1, "x": a["x"], "y": a["y"], "x_grad": a["x_grad"],
        "y_grad": a["y_grad"], "scale": a["weight"], "bias": a["bias"], "prev_mean": a["prev_mean"],
        "prev_var": a["prev_var"], "scale_grad": a["weight_grad"], "bias_grad": a["bias_grad"],
        "tol_y": 6e-4, "tol_x_grad": 7e-4, "tol_param_grad": 6e-4,
    }
    # --- GEMM 3x4 . 4x5 (zenu-matrix/src/operation/mul.rs:307-337)
    src = lines("zenu-matrix/src/operation/mul.rs", 307, 337)
    vs = [[float(x) for x in re.findall(NUM, m)] for m in re.findall(r"vec!\[([^\]]*)\]", src, re.S)]
    lit["gemm_3x4_4x5"] = {"a": vs[0], "b": vs[1], "c": vs[2], "tol_asum": 1e-6}
    # --- ReLU (zenu-matrix/src/operation/relu.rs:255-275)
    lit["relu"] = {"x": [1.0, -1.0, 0.0, 2.0], "y": [1.0, 0.0, 0.0, 2.0], "mask": [1.0, 0.0, 0.0, 1.0], "tol": 1e-6}
    # --- conv fwd literal (zenu-cuda/src/cudnn/graph_conv.rs:433-623)
    a = named_arrays(lines("zenu-cuda/src/cudnn/graph_conv.rs", 433, 580))
    lit["conv_fwd_small"] = {"x_shape": [1, 2, 4, 4], "w_shape": [3, 2, 3, 3], "y_shape": [1, 3, 4, 4],
                             "pad": 1, "stride": 1, "dilation": 1,
                             "input": a["input"], "filter": a["filter"], "output": a["output"], "tol": 1e-6}
    # --- optimizer one-step MLP (zenu-optimizer/tests/net_test.rs)
    src = lines("zenu-optimizer/tests/net_test.rs", 79, 92)
    a = named_arrays(src)
    net = {"w1": a["input_parameters"], "b1": a["input_bias"], "w2": a["output_parameters"], "b2": a["output_bias"],
           "input": [0.1, 0.2], "target": [0.1, 0.2, 0.3, 0.4]}
    def ans(lo, hi):
        s = lines("zenu-optimizer/tests/net_test.rs", lo, hi)
        return [[float(x) for x in re.findall(NUM, m)] for m in re.findall(r"vec!\[([^\]]*)\]", s, re.S)]
    sgd = ans(108, 127)
    adam = ans(144, 167)
    adamw = ans(192, 211)
    lit["optim_net"] = {
        "net": net,
        "sgd": {"lr": 0.9, "steps": 1, "w1": sgd[0], "b1": sgd[1], "w2": sgd[2], "b2": sgd[3], "tol": 1e-4},
        "adam": {"lr": 0.01, "beta1": 0.9, "beta2": 0.999, "eps": 1e-8, "steps": 2, "fresh_state_each_step": True,
                 "w1": adam[0], "b1": adam[1], "w2": adam[2], "b2": adam[3], "tol": 2e-4},
        "adamw": {"lr": 0.01, "beta1": 0.9, "beta2": 0.999, "eps": 1e-8, "weight_decay": 0.01, "steps": 2,
                  "w1": adamw[0], "b1": adamw[1], "w2": adamw[2], "b2": adamw[3], "tol": 2e-4},
    }
    with open(os.path.join(OUT, "literals.json"), "w") as f:
        json.dump(lit, f, indent=1)
    for k, v in lit.items():
        print(k, {kk: (len(vv) if isinstance(vv, list) else vv) for kk, vv in v.items() if not isinstance(vv, dict)})


def c_prototypes(path, text=None):
    """{name: {"ret": type, "args": [types]}} of every function prototype in a C header (comments / preprocessor lines dropped,
    parameter names dropped, whitespace normalised: "float *a" -> "float*").  `text`: already-preprocessed source instead of a file."""
    src = text if text is not None else open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    src = re.sub(r"typedef\s+(?:struct|enum)\s*\w*\s*\{.*?\}\s*\w+\s*;", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b([A-Za-z_]\w*)\s*\(([^()]*)\)\s*;", src):
        ret, name, args = " ".join(m.group(1).split()), m.group(2), m.group(3).strip()
        if not ret or ret.startswith("typedef") or ret in ("extern", "return"):
            continue
        types = []
        for a in ([] if args in ("", "void") else args.split(",")):
            a = " ".join(a.replace("*", " * ").split())
            toks = a.split(" ")
            if len(toks) > 1 and toks[-1] != "*" and re.match(r"[A-Za-z_]\w*$", toks[-1]):
                toks = toks[:-1]   # parameter name
            types.append("".join(t if t == "*" else (" " + t) for t in toks).strip().replace(" *", "*"))
        out[name] = {"ret": ret.replace(" *", "*"), "args": types}
    return out


def ref_symbols():
    """The reference's two native FFI surfaces on the hot path, as the bindgen'd Rust crates see them."""
    ksys = {}
    kdir = os.path.join(REF, "zenu-cuda-kernel-sys", "kernel")
    for h in sorted(os.listdir(kdir)):
        if h.endswith(".h") and h != "kernel.h":
            for k, v in c_prototypes(os.path.join(kdir, h)).items():
                v["header"] = "zenu-cuda-kernel-sys/kernel/" + h
                ksys[k] = v
    fe_path = "zenu-cudnn-frontend-wrapper-sys/cudnn_frontend_wrapper/include/cudnn_frontend_wrapper.h"
    fe = c_prototypes(os.path.join(REF, fe_path))
    for v in fe.values():
        v["header"] = fe_path
    out = {"kernel_sys": ksys, "cudnn_frontend_wrapper": fe}
    with open(os.path.join(OUT, "ref_symbols.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("ref_symbols", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
    ref_symbols()
