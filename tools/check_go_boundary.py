#!/usr/bin/env python
"""Mechanical check of the Go side of the drop-in boundary (no Go toolchain in the image, so nothing
under go-sdr_b200/go has ever been compiled -- this is the next best thing).

    python tools/check_go_boundary.py [--reference /root/reference]

(a) Every EXPORTED top-level identifier (function, method on an exported type, type, var, const) of
    the reference files that INTEGRATION.md section 3 excludes under `-tags sdr.cuda` is declared by the
    twins (go-sdr_b200/go/root, go-sdr_b200/go/stream) with the same signature -- parameter and result
    TYPES compared, names ignored.  Needs the reference tree; skipped (reported) without it.
(b) Every `C.hzsdr_*` the Go code calls is declared in include/hzsdr_cuda.h, and every function the
    header declares is bound somewhere under go-sdr_b200/go.
(c) No identifier that only an excluded file declares is still used by the files that stay in the
    package (or by the twins themselves) without the twins declaring it.  Needs the reference tree.
(d) Every C symbol INTEGRATION.md's table names exists in the header.

Checks (b) and (d) need only this repository: tests/test_go_boundary.py runs them everywhere and (a), (c)
whenever /root/reference is present (the committed expectation list tests/golden/go_boundary_expected.json
pins (a) on boxes without the reference).
"""
from __future__ import annotations

import argparse
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "go-sdr_b200", "go")
HEADER = os.path.join(ROOT, "include", "hzsdr_cuda.h")
INTEGRATION = os.path.join(ROOT, "INTEGRATION.md")

# reference files that get `//go:build !sdr.cuda`, by package
EXCLUDED = {
    "root": ["conv.go", "copy.go"],
    "stream": ["stream/convert.go", "stream/shifter.go", "stream/convolution.go", "stream/decimate.go",
               "stream/downsample.go", "stream/multiply.go", "stream/gain.go", "stream/add.go", "stream/beamform.go"],
    "debug": ["debug/build.go"],
}
TWINS = {"root": [os.path.join(GO, "root")], "stream": [os.path.join(GO, "stream")], "debug": [os.path.join(GO, "debug")]}
REF_PKG_DIR = {"root": "", "stream": "stream", "debug": "debug"}


def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    out = []
    for line in src.splitlines():
        # drop // comments (good enough: the declarations we parse carry no string literals with //)
        i = line.find("//")
        out.append(line if i < 0 else line[:i])
    return "\n".join(out)


def split_top(s: str, sep: str = ",") -> list[str]:
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        parts.append("".join(cur))
    return [p.strip() for p in parts if p.strip()]


def has_top_space(s: str) -> bool:
    depth = 0
    for i, ch in enumerate(s):
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        elif ch == " " and depth == 0:
            # "func (...)" / "chan T" / "map[..] T" start with a keyword, not a name
            head = s[:i]
            if head in ("func", "chan", "map", "struct", "interface", "<-chan"):
                continue
            return True
    return False


def types_of(param_list: str) -> list[str]:
    """Types of a Go parameter / result list, names dropped ("a, b T" -> [T, T])."""
    pieces = split_top(param_list)
    if not pieces:
        return []
    named = any(has_top_space(p) for p in pieces)
    if not named:
        return [norm_type(p) for p in pieces]
    out, pending = [], 0
    for p in pieces:
        if has_top_space(p):
            t = p.split(" ", 1)[1].strip()
            out.extend([norm_type(t)] * (pending + 1))
            pending = 0
        else:
            pending += 1
    return out


def norm_type(t: str) -> str:
    return re.sub(r"\s+", "", t)


def matching_paren(s: str, i: int) -> int:
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses")


def declarations(src: str) -> dict:
    """name -> signature for every top-level declaration.  Functions: 'func(T1,T2)(R1,R2)'; methods are
    keyed 'Recv.Name'; types 'type'; vars / consts 'var' / 'const'."""
    src = strip_comments(src)
    decls = {}
    for m in re.finditer(r"^func\s*", src, flags=re.M):
        i = m.end()
        recv = None
        if src[i] == "(":
            j = matching_paren(src, i)
            r = src[i + 1:j].strip().split()
            recv = r[-1].lstrip("*") if r else None
            i = j + 1
        nm = re.match(r"\s*([A-Za-z_][A-Za-z0-9_]*)\s*", src[i:])
        if not nm:
            continue
        name = nm.group(1)
        i += nm.end()
        if src[i] != "(":
            continue
        j = matching_paren(src, i)
        params = src[i + 1:j]
        rest = src[j + 1:]
        k = rest.find("{")
        results = rest[:k].strip() if k >= 0 else rest.strip()
        if results.startswith("("):
            results = results[1:matching_paren(results, 0)]
        sig = "func(" + ",".join(types_of(params)) + ")(" + ",".join(types_of(results)) + ")"
        decls[(recv + "." + name) if recv else name] = sig
    for m in re.finditer(r"^type\s+([A-Za-z_][A-Za-z0-9_]*)\b", src, flags=re.M):
        decls[m.group(1)] = "type"
    for kw in ("var", "const"):
        for m in re.finditer(r"^%s\s+([A-Za-z_][A-Za-z0-9_]*)\b" % kw, src, flags=re.M):
            decls[m.group(1)] = kw
        for m in re.finditer(r"^%s\s*\((.*?)^\)" % kw, src, flags=re.M | re.S):
            for line in m.group(1).splitlines():
                lm = re.match(r"\s*([A-Za-z_][A-Za-z0-9_]*)\b", line)
                if lm:
                    decls[lm.group(1)] = kw
    return decls


def exported(name: str) -> bool:
    parts = name.split(".")
    return all(p[:1].isupper() for p in parts)


def read(path: str) -> str:
    with open(path, encoding="utf-8") as fh:
        return fh.read()


def go_files(dirs: list[str]) -> list[str]:
    out = []
    for d in dirs:
        for base, _, files in os.walk(d):
            out += [os.path.join(base, f) for f in sorted(files) if f.endswith(".go")]
    return out


def twin_decls(pkg: str) -> dict:
    d = {}
    for f in go_files(TWINS[pkg]):
        d.update(declarations(read(f)))
    return d


def reference_expectations(ref: str) -> dict:
    """{pkg: {exported name: signature}} from the excluded reference files."""
    exp = {}
    for pkg, files in EXCLUDED.items():
        d = {}
        for f in files:
            for name, sig in declarations(read(os.path.join(ref, f))).items():
                if exported(name):
                    d[name] = sig
        exp[pkg] = d
    return exp


def check_a(expect: dict) -> list[str]:
    errs = []
    for pkg, want in expect.items():
        have = twin_decls(pkg)
        for name, sig in sorted(want.items()):
            if name not in have:
                errs.append(f"(a) {pkg}: `{name}` ({sig}) is declared by an excluded reference file but by no twin")
            elif have[name] != sig:
                errs.append(f"(a) {pkg}: `{name}` signature differs: reference {sig}, twin {have[name]}")
    return errs


def header_functions() -> set[str]:
    src = strip_comments(read(HEADER))
    return set(re.findall(r"\b(hzsdr_[a-z0-9_]+)\s*\(", src))


def check_b() -> list[str]:
    hdr = header_functions()
    used = set()
    for f in go_files([GO]):
        used |= set(re.findall(r"\bC\.(hzsdr_[a-z0-9_]+)\s*\(", read(f)))
    errs = [f"(b) Go calls C.{s}, which include/hzsdr_cuda.h does not declare" for s in sorted(used - hdr)]
    errs += [f"(b) include/hzsdr_cuda.h declares {s}, which nothing under go-sdr_b200/go binds" for s in sorted(hdr - used)]
    return errs


def check_c(ref: str) -> list[str]:
    errs = []
    for pkg, files in EXCLUDED.items():
        excluded_paths = {os.path.join(ref, f) for f in files}
        only_excluded = {}
        for p in excluded_paths:
            for name in declarations(read(p)):
                if "." not in name:
                    only_excluded[name] = os.path.relpath(p, ref)
        pkg_dir = os.path.join(ref, REF_PKG_DIR[pkg])
        staying = [os.path.join(pkg_dir, f) for f in sorted(os.listdir(pkg_dir))
                   if f.endswith(".go") and not f.endswith("_test.go") and os.path.join(pkg_dir, f) not in excluded_paths]
        staying_decl = set()
        for p in staying:
            staying_decl |= {n for n in declarations(read(p)) if "." not in n}
        have = twin_decls(pkg)
        users = [(p, strip_comments(read(p))) for p in staying + go_files(TWINS[pkg])]
        for name, origin in sorted(only_excluded.items()):
            if name in staying_decl or name in have:
                continue
            pat = re.compile(r"(?<![.\w])%s\b" % re.escape(name))
            for p, text in users:
                if pat.search(text):
                    errs.append(f"(c) {pkg}: `{name}` (declared only in excluded {origin}) is used by "
                                f"{os.path.relpath(p, ref) if p.startswith(ref) else os.path.relpath(p, ROOT)} and no twin declares it")
                    break
    return errs


def check_d() -> list[str]:
    hdr = header_functions()
    text = read(INTEGRATION)
    named = set()
    for tok in re.findall(r"`(hzsdr_[a-z0-9_/*]+)`", text):
        if "*" in tok:
            prefix = tok.split("*")[0]
            if not any(h.startswith(prefix) for h in hdr):
                named.add(tok)
            continue
        parts = tok.split("/")
        stem = parts[0].rsplit("_", 1)[0] + "_" if len(parts) > 1 else ""
        names = [parts[0]] + [stem + p for p in parts[1:]]
        for n in names:
            if n not in hdr:
                named.add(n)
    return [f"(d) INTEGRATION.md names `{n}`, which include/hzsdr_cuda.h does not declare" for n in sorted(named)]


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--write-expected", action="store_true", help="refresh tests/golden/go_boundary_expected.json from the reference")
    args = ap.parse_args()
    errs = []
    golden = os.path.join(ROOT, "tests", "golden", "go_boundary_expected.json")
    if os.path.isdir(args.reference):
        expect = reference_expectations(args.reference)
        if args.write_expected:
            with open(golden, "w") as fh:
                json.dump(expect, fh, indent=1, sort_keys=True)
                fh.write("\n")
        errs += check_a(expect) + check_c(args.reference)
    else:
        print(f"reference tree {args.reference} not present: (a) from the committed expectation list, (c) skipped")
        with open(golden) as fh:
            errs += check_a(json.load(fh))
    errs += check_b() + check_d()
    for e in errs:
        print(e)
    print(f"{len(errs)} problem(s)")
    return 1 if errs else 0


if __name__ == "__main__":
    sys.exit(main())
