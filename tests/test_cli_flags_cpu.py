"""The drop-in entry points and loader plug-ins expose the reference's command-line surface: every option the reference's
``add_argument`` calls define (tests/golden/cli_flags.json, extracted from the reference sources by tests/golden/make_cli_flags.py)
exists here with the same default, type and action.  Deviations are listed explicitly with their reason."""
import argparse
import json
import os

import pytest

# (entry, option) -> reason.  Everything else must match exactly.
DEVIATIONS = {
    ("train_mbr", "--loader"): "only the on-the-fly loader feeds the trainers here (choices ['otf_utt']); the reference's MBR script lists "
                               "['utt', 'frame', 'otf_utt'] with default 'frame', a module its tree does not contain",
    ("decode", "--loader"): "only the utterance loader exists here (choices ['utt'], default 'utt'); the reference defaults to 'frame', a loader "
                            "module its tree does not contain",
}


def _surface(parser):
    out = {}
    for a in parser._actions:
        if isinstance(a, argparse._HelpAction):
            continue
        name = a.option_strings[0] if a.option_strings else a.dest
        kind = "store_true" if isinstance(a, argparse._StoreTrueAction) else ("store_false" if isinstance(a, argparse._StoreFalseAction) else None)
        out[name] = dict(default=a.default, type=getattr(a.type, "__name__", None) if a.type else None, action=kind,
                         choices=list(a.choices) if a.choices else None)
    return out


def _parsers():
    from pika_b200.decoder import decode_transducer as D
    from pika_b200.loader import otf_utt_loader as OL, utt_loader as UL
    from pika_b200.trainer import train_transducer_bmuf_otfaug as T, train_transducer_mbr_bmuf_otfaug as M
    from pika_b200.utils import compute_global_cmvn as C                                                                   # noqa: F401
    res = {"train": T.build_parser(), "train_mbr": M.build_parser(), "decode": D.build_parser()}
    for key, mod in (("otf_utt_loader", OL), ("utt_loader", UL)):
        p = argparse.ArgumentParser()
        mod.register(p)
        res[key] = p
    return res


@pytest.mark.parametrize("entry", ["train", "train_mbr", "decode", "otf_utt_loader", "utt_loader"])
def test_entry_point_flags_match_reference(golden_dir, entry):
    ref = json.load(open(os.path.join(golden_dir, "cli_flags.json")))[entry]
    have = _surface(_parsers()[entry])
    assert len(ref) >= 10
    for r in ref:
        name = r["names"][0]
        if (entry, name) in DEVIATIONS:
            continue
        assert name in have, "%s: option %s of the reference is missing" % (entry, name)
        h = have[name]
        if "default" in r:
            assert h["default"] == r["default"], (entry, name, h["default"], r["default"])
        if r.get("action") == "store_true":
            assert h["action"] == "store_true" and h["default"] is False, (entry, name)
        if "type" in r:
            assert h["type"] == r["type"], (entry, name, h["type"], r["type"])
        if "choices" in r:
            assert h["choices"] == list(r["choices"]), (entry, name, h["choices"], r["choices"])


def test_cmvn_tool_flags_match_reference(golden_dir):
    """compute_global_cmvn builds its parser inside main(): check the reference's options against its source"""
    import inspect
    from pika_b200.utils import compute_global_cmvn as C
    src = inspect.getsource(C.main)
    for r in json.load(open(os.path.join(golden_dir, "cli_flags.json")))["cmvn"]:
        assert "'%s'" % r["names"][0] in src, r["names"][0]
