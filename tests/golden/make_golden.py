#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the inflate / LZO path.

Reads the OCaml test sources under /root/reference/test (DATA only: string
literals, expected results, line numbers) and writes JSON fixtures next to this
script.  Runs only in the build container (the reference tree does not travel to
the GPU box); the JSON files are committed.

    python tests/golden/make_golden.py [/root/reference]

Outputs
    inflate_ns.json      test/test_ns.ml  De.Inf.Ns.inflate KATs   (src, Ok(consumed, produced)+bytes | Error tag)
    inflate_stream.json  test/test.ml     De.Inf (streaming) KATs  (src, output bytes | Malformed message)
    lzo.json             test/test_lzo.ml + test/test.ml  Lzo.uncompress[_with_buffer] KATs
Vectors whose input is produced by running the reference's *encoder*
(encode_dynamic ...) cannot be extracted without an OCaml toolchain; they are
listed under "skipped" with the reason.
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- lexer
def lex(text):
    """OCaml-ish tokens: ('str', bytes) ('chr', int) ('int', n) ('id', s) ('op', s); each with line."""
    toks = []
    i, n, line = 0, len(text), 1
    while i < n:
        c = text[i]
        if c == "\n":
            line += 1
            i += 1
        elif c in " \t\r":
            i += 1
        elif text.startswith("(*", i):
            depth, i = 1, i + 2
            while i < n and depth:
                if text.startswith("(*", i):
                    depth += 1
                    i += 2
                elif text.startswith("*)", i):
                    depth -= 1
                    i += 2
                else:
                    if text[i] == "\n":
                        line += 1
                    i += 1
        elif c == '"':
            start_line = line
            i += 1
            out = bytearray()
            while text[i] != '"':
                ch = text[i]
                if ch == "\\":
                    nx = text[i + 1]
                    if nx == "\n":  # line continuation: skip newline + leading blanks
                        line += 1
                        i += 2
                        while text[i] in " \t":
                            i += 1
                    elif nx == "x":
                        out.append(int(text[i + 2:i + 4], 16))
                        i += 4
                    elif nx == "o":
                        out.append(int(text[i + 2:i + 5], 8))
                        i += 5
                    elif nx.isdigit():
                        out.append(int(text[i + 1:i + 4]))
                        i += 4
                    else:
                        out.append({"n": 10, "t": 9, "b": 8, "r": 13, " ": 32, "\\": 92, '"': 34, "'": 39}[nx])
                        i += 2
                else:
                    if ch == "\n":
                        line += 1
                    out += ch.encode("latin-1") if ord(ch) < 256 else ch.encode("utf-8")
                    i += 1
            i += 1
            toks.append(("str", bytes(out), start_line))
        elif c == "'" and re.match(r"'(\\x[0-9a-fA-F]{2}|\\[0-9]{3}|\\.|[^\\'])'", text[i:]):
            m = re.match(r"'(\\x[0-9a-fA-F]{2}|\\[0-9]{3}|\\.|[^\\'])'", text[i:])
            body = m.group(1)
            if body.startswith("\\x"):
                v = int(body[2:], 16)
            elif body.startswith("\\") and body[1:].isdigit():
                v = int(body[1:])
            elif body.startswith("\\"):
                v = {"n": 10, "t": 9, "b": 8, "r": 13, " ": 32, "\\": 92, '"': 34, "'": 39}[body[1]]
            else:
                v = ord(body)
            toks.append(("chr", v, line))
            i += m.end()
        elif c.isdigit():
            m = re.match(r"0[xX][0-9a-fA-F_]+|[0-9][0-9_]*", text[i:])
            toks.append(("int", int(m.group(0).replace("_", ""), 0), line))
            i += m.end()
        elif c.isalpha() or c == "_" or c == "`":
            m = re.match(r"`?[A-Za-z_][A-Za-z0-9_'.]*", text[i:])
            toks.append(("id", m.group(0), line))
            i += m.end()
        else:
            m = re.match(r"@@|->|::|[\^\[\]\(\);,=\-\+\*<>|!~:#&{}?/.@$%]", text[i:])
            toks.append(("op", m.group(0), line))
            i += m.end()
    return toks


def split_defs(text):
    """Top-level `let name` definitions -> list of (name, first_line, tokens, raw_text)."""
    lines = text.split("\n")
    starts = [k for k, l in enumerate(lines) if re.match(r"let\s+(rec\s+)?[a-z_][A-Za-z0-9_']*", l)]
    defs = []
    for a, b in zip(starts, starts[1:] + [len(lines)]):
        raw = "\n".join(lines[a:b])
        name = re.match(r"let\s+(?:rec\s+)?([a-z_][A-Za-z0-9_']*)", lines[a]).group(1)
        toks = [(k, v, ln + a) for (k, v, ln) in lex(raw)]
        defs.append((name, a + 1, toks, raw))
    return defs


# ----------------------------------------------------------------------- evaluator
class NoEval(Exception):
    pass


def ev(toks, pos, env):
    """Evaluate a string-valued expression starting at toks[pos]; returns (value, next_pos)."""
    val, pos = ev_atom(toks, pos, env)
    while pos < len(toks) and toks[pos][:2] == ("op", "^"):
        rhs, pos = ev_atom(toks, pos + 1, env)
        val = val + rhs
    return val, pos


def ev_atom(toks, pos, env):
    k, v, _ = toks[pos]
    if k == "str":
        return v, pos + 1
    if (k, v) == ("op", "("):
        val, pos = ev(toks, pos + 1, env)
        if toks[pos][:2] != ("op", ")"):
            raise NoEval("paren")
        return val, pos + 1
    if (k, v) == ("op", "["):
        pos += 1
        out = b""
        while toks[pos][:2] != ("op", "]"):
            item, pos = ev(toks, pos, env)
            out += item
            if toks[pos][:2] == ("op", ";"):
                pos += 1
        return out, pos + 1  # lists are only ever String.concat ""-ed
    if k == "id":
        if v == "String.concat":
            if toks[pos + 1][:2] != ("str", b""):
                raise NoEval("concat sep")
            return ev_atom(toks, pos + 2, env)
        if v == "String.make":
            n, c = toks[pos + 1], toks[pos + 2]
            if n[0] != "int" or c[0] != "chr":
                raise NoEval("String.make")
            return bytes([c[1]]) * n[1], pos + 3
        if v in ("bigstring_of_string", "Bstr.of_string"):
            return ev_atom(toks, pos + 1, env)
        if v in env:
            return env[v], pos + 1
    raise NoEval("token %r" % (toks[pos],))


def bindings(toks, env=None):
    """Collect `let x = <expr> in` bindings whose rhs evaluates to bytes (later bindings shadow)."""
    env = dict(env or {})
    lines = {}
    for p in range(len(toks) - 3):
        if toks[p][:2] == ("id", "let") and toks[p + 1][0] == "id" and toks[p + 2][:2] == ("op", "="):
            try:
                val, q = ev(toks, p + 3, env)
            except (NoEval, IndexError, KeyError):
                continue
            if q < len(toks) and toks[q][:2] == ("id", "in"):
                env[toks[p + 1][1]] = val
                lines[toks[p + 1][1]] = toks[p + 1][2]
    return env, lines


def title_of(raw):
    m = re.search(r'Alcotest\.test_case\s+"([^"]*)"', raw)
    return m.group(1) if m else None


# ------------------------------------------------------------------------ recipes
def extract_ns(path):
    text = open(path, encoding="latin-1").read()
    out, skipped = [], []
    for name, line, toks, raw in split_defs(text):
        if "Inf.Ns.inflate" not in raw or title_of(raw) is None:
            continue
        env, _ = bindings(toks)
        cite = "test/test_ns.ml:%d" % line
        if "src" not in env:
            skipped.append({"name": name, "cite": cite,
                            "reason": "input is produced by the reference encoder (Def.encode); needs OCaml"})
            continue
        src = env["src"]
        m_dst = re.search(r"let dst = bigstring_create (\d+)", raw)
        dst_len = int(m_dst.group(1)) if m_dst else None  # None = module-level dst (large)
        m_err = re.search(r"\(Error `([A-Za-z_]+)\)", raw)
        ent = {"name": name, "title": title_of(raw), "cite": cite, "src": src.hex(), "dst_len": dst_len}
        if m_err:
            ent["error"] = m_err.group(1)
        else:
            m_ok = re.search(r"\(Ok\s*\(\s*([^,]+),\s*([^)]*)\)\)", raw)
            if not m_ok:
                skipped.append({"name": name, "cite": cite, "reason": "no Ok/Error pattern"})
                continue
            cons = m_ok.group(1).strip()
            mm = re.match(r"De\.bigstring_length src(?:\s*-\s*(\d+))?$", cons)
            consumed = len(src) - int(mm.group(1) or 0) if mm else int(cons)
            prod = m_ok.group(2).strip()
            if prod.startswith("String.length expected"):
                produced = len(env["expected"])
                ent["out"] = env["expected"].hex()
            else:
                produced = int(prod)
                if "expected" in env:
                    ent["out"] = env["expected"].hex()
            ent["ok"] = [consumed, produced]
        out.append(ent)
    return out, skipped


def extract_stream(path):
    text = open(path, encoding="latin-1").read()
    out, skipped = [], []
    for name, line, toks, raw in split_defs(text):
        if "Inf.decoder" not in raw or title_of(raw) is None:
            continue
        cite = "test/test.ml:%d" % line
        env, _ = bindings(toks)
        # input: Inf.decoder (`String EXPR)
        src = None
        for p, t in enumerate(toks):
            if t[:2] == ("id", "Inf.decoder") and toks[p + 1][:2] == ("op", "(") and toks[p + 2][:2] == ("id", "`String"):
                try:
                    src, _ = ev(toks, p + 3, env)
                except (NoEval, IndexError):
                    src = None
                break
        if src is None:
            skipped.append({"name": name, "cite": cite, "reason": "input not a literal (`Manual feed or encoder output)"})
            continue
        ent = {"name": name, "title": title_of(raw), "cite": cite, "src": src.hex()}
        m_err = re.search(r'`Malformed\s+"([^"]*)"', raw)
        if m_err:
            ent["malformed"] = m_err.group(1)
            if '"non-corrupted output" ""' in raw:
                ent["out"] = ""
            out.append(ent)
            continue
        # expected output: Alcotest.(check string) "title" EXPR ...
        got = None
        for p, t in enumerate(toks):
            if (t[:2] == ("id", "check") and toks[p + 1][:2] in (("id", "string"), ("id", "str"))
                    and toks[p + 2][:2] == ("op", ")") and toks[p + 3][0] == "str"):
                q = p + 5 if toks[p + 4][:2] == ("id", "res") else p + 4
                try:
                    got, _ = ev(toks, q, env)
                except (NoEval, IndexError):
                    got = None
                break
        if got is None:
            skipped.append({"name": name, "cite": cite, "reason": "expected output not a literal"})
            continue
        ent["out"] = got.hex()
        out.append(ent)
    return out, skipped


def extract_lzo(paths):
    out, skipped = [], []
    for path, tag in paths:
        text = open(path, encoding="latin-1").read()
        genv = {}
        defs = split_defs(text)
        for name, line, toks, raw in defs:  # module-level string tables (output_lzo_3, ...)
            if title_of(raw) is None and not raw.startswith("let () ="):
                try:
                    p = next(k for k, t in enumerate(toks) if t[:2] == ("op", "="))
                    val, _ = ev(toks, p + 1, genv)
                    genv[name] = val
                except (NoEval, IndexError, StopIteration, KeyError):
                    pass
        for name, line, toks, raw in defs:
            if title_of(raw) is None or "Lzo.uncompress" not in raw:
                continue
            cite = "%s:%d" % (tag, line)
            if "Lzo.compress" in raw:
                skipped.append({"name": name, "cite": cite, "reason": "input is the output of the reference compressor; needs OCaml"})
                continue
            env, _ = bindings(toks, genv)
            api = "uncompress_with_buffer" if "Lzo.uncompress_with_buffer" in raw else "uncompress"
            # shape A: several `Alcotest.(check res) "title" (Ok EXPR | Error ..) (uncompress EXPR)` in one test
            multi = []
            for p, t in enumerate(toks):
                if not (t[:2] == ("id", "check") and toks[p + 1][:2] == ("id", "res") and toks[p + 2][:2] == ("op", ")")
                        and toks[p + 3][0] == "str" and toks[p + 4][:2] == ("op", "(")):
                    continue
                try:
                    if toks[p + 5][:2] == ("id", "Ok"):
                        want, q = ev_atom(toks, p + 6, env)
                        err = None
                    elif toks[p + 5][:2] == ("id", "Error"):
                        q = p + 6
                        depth = 0
                        err = None
                        while True:  # skip the parenthesised error value, remember its tag
                            if toks[q][:2] == ("op", "("):
                                depth += 1
                            elif toks[q][:2] == ("op", ")"):
                                depth -= 1
                                if depth == 0:
                                    q += 1
                                    break
                            elif toks[q][0] == "id" and toks[q][1].startswith("`"):
                                err = toks[q][1][1:]
                            q += 1
                        want = None
                    else:
                        continue
                    if toks[q][:2] != ("op", ")") or toks[q + 1][:2] != ("op", "(") or toks[q + 2][:2] != ("id", "uncompress"):
                        continue
                    src, _ = ev_atom(toks, q + 3, env)
                except (NoEval, IndexError):
                    continue
                if len(src) > 200000:
                    continue  # the >1 MiB zero runs are rebuilt compactly in manual_lzo
                ent = {"name": "%s_%d" % (name, len(multi)), "title": toks[p + 3][1].decode("latin-1"), "cite": cite,
                       "api": api, "src": src.hex(), "raw_hint": "ok" if err is None else "error"}
                if want is not None:
                    ent["out"] = want.hex()
                if err:
                    ent["error"] = err
                multi.append(ent)
            if multi:
                out += multi
                continue
            # shape B: one input binding, one expected binding / literal
            src = None
            for key in ("input", "src", "compressed"):
                if key in env:
                    src = env[key]
                    break
            if src is None:
                skipped.append({"name": name, "cite": cite, "reason": "irregular shape (see manual_lzo)"})
                continue
            ent = {"name": name, "title": title_of(raw), "cite": cite, "api": api, "src": src.hex()}
            exp = None
            for key in ("expect", "expected", "output"):
                if key in env:
                    exp = env[key]
                    break
            if exp is None:
                for p, t in enumerate(toks):
                    if (t[:2] == ("id", "check") and toks[p + 1][:2] in (("id", "str"), ("id", "string"))
                            and toks[p + 2][:2] == ("op", ")") and toks[p + 3][0] == "str"):
                        for q in (p + 4, p + 5):
                            try:
                                exp, _ = ev(toks, q, env)
                                break
                            except (NoEval, IndexError):
                                continue
                        break
            if exp is not None:
                ent["out"] = exp.hex()
                ent["raw_hint"] = "ok"
            else:
                ent["raw_hint"] = "error" if re.search(r"Ok _ -> Alcotest.fail", raw) else "ok"
            out.append(ent)
    return out, skipped


def manual_lzo(path):
    """Vectors whose OCaml shape is irregular: rebuilt here, each literal checked to be present in the source."""
    text = open(path, encoding="latin-1").read()
    defs = {name: (line, toks) for name, line, toks, raw in split_defs(text)}

    def lits(name):
        return [v for k, v, _ in defs[name][1] if k == "str"], [v for k, v, _ in defs[name][1] if k == "int"]

    out = []
    s2, _ = lits("test_lzo_2")
    assert b"\x00\x00\x11\x00\x00" in s2
    out.append({"name": "test_lzo_2", "title": "ensure it fails without exception", "api": "uncompress_with_buffer",
                "cite": "test/test_lzo.ml:%d" % defs["test_lzo_2"][0], "src": "0000110000", "raw_hint": "error"})
    s4, _ = lits("test_lzo_4")
    assert b"\x11\x00\x00" in s4
    out.append({"name": "test_lzo_4", "title": "handle empty", "api": "uncompress_with_buffer",
                "cite": "test/test_lzo.ml:%d" % defs["test_lzo_4"][0], "src": "110000", "out": "", "raw_hint": "ok"})
    s15, i15 = lits("test_lzo_15")
    for lit in (b"\x00\xb0", b"\x00\x00\x7f", b"\x11\x00\x00", b"\x01", b"\x00", b"\x01\x11\x00\x00"):
        assert lit in s15, lit
    for n in (176, 382, 33151, 130, 76501, 300, 4210753):
        assert n in i15, n
    term = b"\x11\x00\x00"
    cite = "test/test_lzo.ml:%d" % defs["test_lzo_15"][0]
    cases = [(b"\x00\xb0", 176), (b"\x00\x00\x7f", 382), (b"\x00" + b"\x00" * 130 + b"\x01", 33151),
             (b"\x00" + b"\x00" * 300 + b"\x01", 76501)]
    for k, (prefix, extra) in enumerate(cases):
        body = b"A" * (3 + 15 + extra)
        out.append({"name": "test_lzo_15_manual_%d" % k, "title": "CVE 2017 8845 literal run 3+15+%d" % extra,
                    "api": "uncompress_with_buffer", "cite": cite, "src": (prefix + body + term).hex(),
                    "out_repeat": ["41", len(body)], "raw_hint": "ok"})
    out.append({"name": "test_lzo_15_manual_4", "title": "CVE 2017 8845 literal run overflows (2**30)-1",
                "api": "uncompress_with_buffer", "cite": cite,
                "src_repeat": ["00", 1 + 4210753, "01110000"], "raw_hint": "error", "error": "Malformed"})
    return out


def extract_encoder(path):
    """The encoder-side known answers of test/test.ml: huffman_length_extra (507-531: commands -> 16 bytes), tree_0
    (1169-1190), the rfc5322 histogram and the two (length, code) pairs it pins (1192-1237), the level-0 Ns block
    (test_ns.ml:1189-1222 is in inflate_ns.json's neighbourhood; its bytes are pinned in tests/test_gpu_deflate.py)."""
    text = open(path).read()
    lines = text.split("\n")

    def line_of(needle):
        return next(i + 1 for i, l in enumerate(lines) if needle in l)

    out = {}
    a = text.index('Alcotest.test_case "huffman length extra"')
    blk = text[a:text.index("let fuzz10 ()", a)]
    m = re.search(r'"encoding" res "((?:[^"\\]|\\.)*)"', blk)
    raw = m.group(1)
    # OCaml string escapes: \ddd is DECIMAL
    b = bytearray()
    i = 0
    while i < len(raw):
        if raw[i] == "\\" and raw[i + 1].isdigit():
            b.append(int(raw[i + 1:i + 4])); i += 4
        elif raw[i] == "\\":
            b.append({"n": 10, "t": 9, "\\": 92, '"': 34}[raw[i + 1]]); i += 2
        else:
            b.append(ord(raw[i])); i += 1
    assert "`Literal '\\000'; `Literal '\\000'; `Copy (1, 258); `Copy (1, 256)" in " ".join(blk.split())
    copy = lambda off, ln: 0x2000000 | ((ln - 3) << 16) | (off - 1)   # Queue.cmd, lib/de.ml:2245-2266
    out["huffman_length_extra"] = {"cite": "test/test.ml:%d" % line_of('"huffman length extra"'),
                                   "cmds": [0, 0, copy(1, 258), copy(1, 256), 256], "last": True, "bytes": bytes(b).hex()}
    a = text.index('Alcotest.test_case "rfc5322 corpus literals/lengths freqs"')
    blk = text[a:text.index("let compress_and_uncompress", a)]
    arr = blk[blk.index("[|") + 2:blk.index("|]")]
    freqs = [int(x) for x in re.split(r"[;\s]+", arr.strip()) if x]
    assert len(freqs) == 2 * 286 + 1
    m6 = re.search(r'"literal 6" \((\d+), (0x[0-9a-f]+)\)', blk)
    me = re.search(r'"eob" \((\d+), (0x[0-9a-f]+)\)', blk)
    out["tree_rfc5322"] = {"cite": "test/test.ml:%d" % line_of('"rfc5322 corpus literals/lengths freqs"'), "freqs": freqs,
                           "literal_6": [int(m6.group(1)), int(m6.group(2), 16)], "eob": [int(me.group(1)), int(me.group(2), 16)]}
    out["tree_0"] = {"cite": "test/test.ml:%d" % line_of('"empty literals/lengths freqs"'), "lengths": [[0, 1], [256, 1]],
                     "codes": [[0, 0], [256, 1]]}
    return out


def main():
    ns, ns_skip = extract_ns(os.path.join(REF, "test/test_ns.ml"))
    st, st_skip = extract_stream(os.path.join(REF, "test/test.ml"))
    lz, lz_skip = extract_lzo([(os.path.join(REF, "test/test_lzo.ml"), "test/test_lzo.ml"),
                               (os.path.join(REF, "test/test.ml"), "test/test.ml")])
    lz_manual = manual_lzo(os.path.join(REF, "test/test_lzo.ml"))
    lz += lz_manual
    have = {v["name"] for v in lz}
    lz_skip = [x for x in lz_skip if not any(h.startswith(x["name"]) for h in have)]
    for fn, vec, sk in (("inflate_ns.json", ns, ns_skip), ("inflate_stream.json", st, st_skip),
                        ("lzo.json", lz, lz_skip)):
        with open(os.path.join(HERE, fn), "w") as f:
            json.dump({"source": "mirage/decompress v1.6.0 (1b61f758), extracted by tests/golden/make_golden.py",
                       "vectors": vec, "skipped": sk}, f, indent=1)
        print("%s: %d vectors, %d skipped" % (fn, len(vec), len(sk)))
        for s in sk:
            print("   skipped %-34s %s (%s)" % (s["name"], s["cite"], s["reason"]))
    enc = extract_encoder(os.path.join(REF, "test/test.ml"))
    with open(os.path.join(HERE, "encoder.json"), "w") as f:
        json.dump({"source": "mirage/decompress v1.6.0 (1b61f758), extracted by tests/golden/make_golden.py", "kats": enc}, f, indent=1)
    print("encoder.json: %d known answers" % len(enc))


if __name__ == "__main__":
    main()
