"""The reference's command-line surface, extracted by executing its own argparse block (vae/main.py, from
`parser = argparse.ArgumentParser()` to the line before `args = parser.parse_args()`), plus what the README commands parse to.
Writes tests/golden/reference_cli.json; tests/test_host.py checks splitvae_b200.main's parser against it.

    python scripts/make_reference_cli_golden.py          # needs /root/reference (build container only)
"""
import argparse
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/vae/main.py"
README = "/root/reference/README.md"


def main():
    text = open(REF).read().split("\n")
    a = next(i for i, l in enumerate(text) if l.startswith("parser = argparse.ArgumentParser()"))
    b = next(i for i, l in enumerate(text) if l.startswith("args = parser.parse_args()"))
    ns = {"argparse": argparse}
    exec("\n".join(text[a:b]), ns)                                  # vae/main.py:15-31, verbatim
    parser = ns["parser"]
    flags = {}
    for act in parser._actions:
        if act.dest == "help":
            continue
        flags[act.dest] = {"options": act.option_strings, "default": act.default,
                           "type": getattr(act.type, "__name__", None), "store_true": isinstance(act, argparse._StoreTrueAction)}
    # every `python main.py ...` command line of the README, parsed by the reference's own parser
    cmds = []
    for line in open(README).read().split("\n"):
        m = re.search(r"python main\.py(.*)", line)
        if m and "spair" not in line.lower():
            argv = m.group(1).replace("`", "").split()
            try:
                ns_args = parser.parse_args(argv)
            except SystemExit:
                continue
            cmds.append({"argv": argv, "parsed": vars(ns_args)})
    path = os.path.join(ROOT, "tests", "golden", "reference_cli.json")
    with open(path, "w") as f:
        json.dump({"source": f"{REF} lines {a + 1}-{b} executed verbatim; command lines from {README}", "flags": flags, "commands": cmds}, f, indent=1)
    print("wrote", path, len(flags), "flags,", len(cmds), "README commands")


if __name__ == "__main__":
    main()
