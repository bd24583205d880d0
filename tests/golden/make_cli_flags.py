#!/usr/bin/env python
"""Record the command-line surface of the reference's entry points and loader plug-ins as a fixture: for every ``add_argument`` call
in the reference SOURCE (parsed with ``ast``; the files are not imported -- their parsers live under ``__main__`` and pull in PyKaldi)
the option strings, default, type name, action and choices.  Run in the build container only:

    python tests/golden/make_cli_flags.py      ->  tests/golden/cli_flags.json
"""
import ast
import json
import os

REF = "/root/reference"
FILES = {
    "train": "trainer/train_transducer_bmuf_otfaug.py",
    "train_mbr": "trainer/train_transducer_mbr_bmuf_otfaug.py",
    "decode": "decoder/decode_transducer.py",
    "cmvn": "utils/compute_global_cmvn.py",
    "otf_utt_loader": "loader/otf_utt_loader.py",
    "utt_loader": "loader/utt_loader.py",
}


def literal(node):
    try:
        return ast.literal_eval(node)
    except Exception:
        if isinstance(node, ast.Name):
            return node.id                           # type=int -> "int"
        if isinstance(node, ast.BinOp):
            try:
                return eval(compile(ast.Expression(node), "<ast>", "eval"), {})    # 128*1024
            except Exception:
                pass
        return ast.unparse(node)


def flags_of(path):
    tree = ast.parse(open(path).read())
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "add_argument":
            names = [literal(a) for a in node.args]
            kw = {k.arg: literal(k.value) for k in node.keywords if k.arg in ("default", "type", "action", "choices", "nargs")}
            out.append({"names": names, **kw})
    out.sort(key=lambda d: d["names"][0])
    return out


if __name__ == "__main__":
    res = {k: flags_of(os.path.join(REF, v)) for k, v in FILES.items()}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cli_flags.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print({k: len(v) for k, v in res.items()})
