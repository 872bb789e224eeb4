#!/usr/bin/env python
"""Extracts the golden L2/Linf vectors this repository is pinned against from the reference's own test-suite.

Every `Elixir` in tests/elixirs.py cites the `@test_trixi_include` block it mirrors as `test/<file>.jl:<from>-<to>`.
This script reads exactly those lines of the reference checkout (default /root/reference, read-only), parses the
`l2 = [...]` and `linf = [...]` arrays and writes them to tests/golden/reference_goldens.json.  The reference is a
Julia package and cannot travel to the GPU box, the JSON fixture can; tests/test_oracle_golden.py asserts that the
values hard-coded in tests/elixirs.py are bit-for-bit the ones extracted here.

    python tests/golden/extract_reference_goldens.py [/path/to/Trixi.jl]
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def parse_block(text, key):
    m = re.search(r"\b" + key + r"\s*=\s*\[(.*?)\]", text, re.S)
    if not m:
        raise ValueError(f"no {key} array in block")
    return [float(tok) for tok in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", m.group(1))]


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    from elixirs import ELIXIRS
    out = {}
    for name, ex in sorted(ELIXIRS.items()):
        m = re.match(r"(\S+?):(\d+)-(\d+)", ex.source)
        path, lo, hi = m.group(1), int(m.group(2)), int(m.group(3))
        with open(os.path.join(ref, path)) as f:
            lines = f.readlines()
        block = "".join(lines[lo - 1:hi])
        elixir = re.search(r'"(elixir_[a-z0-9_]+\.jl)"', block)
        out[name] = {"source": ex.source, "elixir": elixir.group(1) if elixir else None,
                     "l2": parse_block(block, "l2"), "linf": parse_block(block, "linf")}
        assert out[name]["l2"] == list(ex.l2) and out[name]["linf"] == list(ex.linf), \
            f"{name}: tests/elixirs.py differs from {ex.source}"
    with open(os.path.join(HERE, "reference_goldens.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"{len(out)} golden vectors written")


if __name__ == "__main__":
    main()
