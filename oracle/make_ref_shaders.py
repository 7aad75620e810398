#!/usr/bin/env python
"""make_ref_shaders.py — TEST INFRASTRUCTURE.  Wrap the reference's own GLSL, read where it lies under /root/reference,
so that g++ can compile it against oracle/glsl_shim.h.  Output goes to oracle/_ref/gen/*.inc (git-ignored; reference
text never enters the repository).  Each .inc is the body of one C++ struct: GLSL globals become members, uniforms
become static members, functions become member functions.  The rewrite is purely lexical and touches no arithmetic:

  * `#version` / `#extension` lines dropped; `#include "x"` spliced in from Shader/ (glslc -I Shader)
  * `layout(...) in|out T name;`            -> `T name;`
  * `layout(...) uniform T name;`           -> `inline static T name;`            (sampler, texture2D, uimage3D)
  * `layout(...) uniform Block { ... };`    -> each member `inline static`
  * `layout(...) uniform Block { ... } v;`  -> `struct Block_t { ... }; inline static Block_t v;`
  * floating literals get an `f` suffix (GLSL literals are fp32, C++ literals are double)
  * r-value swizzles `.xyz` -> `.xyz()`; swizzle assignment `a.xy = e;` -> `a.set_xy(e);`
  * `inout T x` / `out T x` parameters -> `T& x`; `const int N = k;` -> `static constexpr`; `T a[n] = T[](...)` -> `{...}`; `uint(` / `int(` -> `to_uint(` / `to_int(` (saturating / NaN-safe conversions of the shim)
  * `discard` -> `throw glsl_discard()`
  * PINNED operand order: in `acc += a * f(..)` where f has an inout parameter, `a` is read before the call
    (GLSL leaves the order open); the rewrite materialises `a` first.

The voxel-pass stages live in Pipelang/Internal/main.lua as `Code [[ ... ]]` strings with `Input`/`Output` declarations
in Lua; those declarations are turned into members and the code strings into `void main()`.

usage: make_ref_shaders.py <reference root> <output dir>
"""
import os
import re
import sys

FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)([fF]?)(?![\w.])")
SWZ = r"(?:[xyzw]{2,4}|[rgba]{2,4}|[stpq]{2,4})"
SWZ_ASSIGN = re.compile(r"([A-Za-z_][\w\[\]\.]*?)\.(xy|st|xz|yz)\s*=(?!=)\s*([^;]+);")
SWZ_ADD_ASSIGN = re.compile(r"([A-Za-z_][\w\[\]\.]*?)\.(st|xy|rgb|xyz)\s*\+=\s*([^;]+);")
SWZ_BAD_ASSIGN = re.compile(r"\.(" + SWZ + r")\s*[-+*/]?=(?!=)")
SWZ_RVALUE = re.compile(r"\.(" + SWZ + r")\b(?!\s*\()")
LAYOUT = r"layout\s*\([^)]*\)\s*"


def splice_includes(path, shader_root, seen=None):
    text = open(path).read()
    def repl(m):
        inc = m.group(1)
        for base in (os.path.dirname(path), shader_root):
            p = os.path.join(base, inc)
            if os.path.exists(p):
                return splice_includes(p, shader_root) + "\n"
        raise SystemExit(f"include not found: {inc} (from {path})")
    return re.sub(r'^[ \t]*#include\s+"([^"]+)"[ \t]*$', repl, text, flags=re.M)


def array_constructors(text):
    """`T name[n] = T [] ( a, b, ... );`  ->  `T name[n] = { a, b, ... };`"""
    out, pos = [], 0
    for m in re.finditer(r"=\s*\w+\s*\[\s*\]\s*\(", text):
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        out.append(text[pos:m.start()] + "= {" + text[m.end():i - 1] + "}")
        pos = i
    return "".join(out) + text[pos:]


def lexical(text, inout_fns=()):
    """the arithmetic-neutral rewrites shared by .frag files and main.lua code strings"""
    text = re.sub(r"^[ \t]*#(version|extension)[^\n]*$", "", text, flags=re.M)
    text = FLOAT_LIT.sub(lambda m: m.group(1) + (m.group(2) or "f"), text)
    text = SWZ_ADD_ASSIGN.sub(lambda m: f"{m.group(1)}.add_{m.group(2)}({m.group(3)});", text)
    text = SWZ_ASSIGN.sub(lambda m: f"{m.group(1)}.set_{m.group(2)}({m.group(3)});", text)
    bad = SWZ_BAD_ASSIGN.search(text)
    if bad:
        raise SystemExit(f"unhandled swizzle assignment near: {text[max(0, bad.start() - 40):bad.end() + 40]!r}")
    text = SWZ_RVALUE.sub(lambda m: f".{m.group(1)}()", text)
    text = re.sub(r"\binout\s+(\w+)\s+(\w+)", r"\1& \2", text)
    text = re.sub(r"([(,]\s*)out\s+(\w+)\s+(\w+)", r"\1\2& \3", text)        # parameter qualifiers
    text = re.sub(r"([(,]\s*)in\s+(\w+)\s+(\w+)", r"\1\2 \3", text)
    text = re.sub(r"^([ \t]*)const\s+int\s+(\w+)\s*=\s*(\d+)\s*;", r"\1static constexpr int \2 = \3;", text, flags=re.M)   # usable as array bounds
    text = array_constructors(text)
    text = re.sub(r"\buint\s*\(", "to_uint(", text)
    text = re.sub(r"(?<![\w.])int\s*\((?!\s*\))", "to_int(", text)
    text = re.sub(r"\bdiscard\b", "throw glsl_discard()", text)
    for fn in inout_fns:
        # acc += a * fn(...);   ->   { auto l_ = a; acc += l_ * fn(...); }
        pat = re.compile(r"^([ \t]*)(\w+)\s*\+=\s*([\w.]+)\s*\*\s*(" + fn + r"\s*\(.*\))\s*;[ \t]*$", re.M)
        text = pat.sub(lambda m: f"{m.group(1)}{{ auto l_ = {m.group(3)}; {m.group(2)} += l_ * {m.group(4)}; }}", text)
    return text


def declarations(text):
    """layout-qualified globals -> members"""
    def block(m):
        name, body, inst = m.group(1), m.group(2), m.group(3)
        if inst:
            return f"struct {name}_t {{{body}}}; inline static {name}_t {inst};"
        members = re.sub(r"^([ \t]*)(\w+)\s+(\w+)\s*;", r"\1inline static \2 \3;", body, flags=re.M)
        return f"/* uniform {name} */{members}"
    text = re.sub(LAYOUT + r"uniform\s+(\w+)\s*\{([^}]*)\}\s*(\w*)\s*;", block, text)
    text = re.sub(LAYOUT + r"uniform\s+(?:readonly\s+|writeonly\s+)?(\w+)\s+(\w+)\s*;", r"inline static \1 \2;", text)
    text = re.sub(LAYOUT + r"(?:flat\s+)?(?:in|out)\s+(\w+)\s+(\w+)\s*;", r"\1 \2;", text)
    if re.search(r"\blayout\s*\(", text):
        raise SystemExit("unhandled layout() declaration: " + re.search(r"[^\n]*\blayout\s*\([^\n]*", text).group(0))
    return text


def undefs(text):
    return "".join(f"#undef {m}\n" for m in re.findall(r"^[ \t]*#define\s+(\w+)", text, flags=re.M))


def frag(ref, rel, inout_fns=()):
    shader_root = os.path.join(ref, "Shader")
    text = splice_includes(os.path.join(shader_root, rel), shader_root)
    fns = set(re.findall(r"\b(\w+)\s*\([^()]*\binout\b[^()]*\)", text)) | set(inout_fns)
    text = declarations(lexical(text, sorted(fns)))
    return f"// generated from {rel} (+ includes) by oracle/make_ref_shaders.py — do not commit\n" + text + "\n" + undefs(text)


# ---------------------------------------------------------------------------------------------- main.lua
def lua_function(lua, name):
    m = re.search(r"^function\s+" + name + r"\s*\(\)(.*?)^end\b", lua, flags=re.M | re.S)
    if not m:
        raise SystemExit(f"main.lua: function {name} not found")
    body = m.group(1)
    ins = re.findall(r'Input\s+"(\w+)"\s+"(\w+)"', body)
    outs = re.findall(r'Output\s+"(\w+)"\s+"(\w+)"', body)
    code = re.search(r"Code\s*\[\[(.*?)\]\]", body, flags=re.S).group(1)
    return ins, outs, code


def lua_uniform_blocks(lua):
    out = []
    for name, members in re.findall(r'Output\s+"uniform"\s+"(\w+)"\s*\[\[(.*?)\]\]', lua, flags=re.S):
        out.append(f"/* uniform {name} */" + re.sub(r"^([ \t]*)(\w+)\s+(\w+)\s*;", r"\1inline static \2 \3;", members, flags=re.M))
    return "\n".join(out)


def voxel_stages(ref):
    lua = open(os.path.join(ref, "Pipelang", "Internal", "main.lua")).read()
    blocks = lua_uniform_blocks(lua)
    # geometry stage: inputs are per-vertex arrays of 3 (InputPrimitive = "triangles", main.lua:77-81)
    gi, go, gcode = lua_function(lua, "VoxelGS")
    gs = ["// generated from Pipelang/Internal/main.lua (VoxelGS) by oracle/make_ref_shaders.py — do not commit", blocks]
    gs += [f"{t} {n}[3];" for t, n in gi if t != "uniform"]
    gs += [f"{t} {n};" for t, n in go]
    gs.append("void main() {" + lexical(gcode) + "}")
    # pixel stage: material function, then the voxel store (Pipelang chains them; outputs of one are inputs of the next)
    mi, mo, mcode = lua_function(lua, "BasicMaterial")
    pi, po, pcode = lua_function(lua, "VoxelPS")
    seen, decl = set(), []
    for t, n in mi + mo + pi + po:
        if n in seen or t == "uniform":
            continue
        seen.add(n)
        decl.append(f"inline static {t} {n};" if t in ("texture2D", "uimage3D", "sampler") else f"{t} {n};")
    samplers = [f"inline static sampler {n};" for n in re.findall(r'Output\s+"sampler"\s+"(\w+)"', lua)]
    ps = ["// generated from Pipelang/Internal/main.lua (BasicMaterial + VoxelPS) by oracle/make_ref_shaders.py — do not commit",
          blocks] + samplers + decl
    ps.append("void main() {\n{" + lexical(mcode) + "}\n{" + lexical(pcode) + "}\n}")
    return "\n".join(gs) + "\n", "\n".join(ps) + "\n"


def main():
    ref, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    files = {
        "indirect_frag.inc": frag(ref, "Lighting/indirect.frag"),
        "gtao_frag.inc": frag(ref, "GTAO/gtao.frag"),
        "gtao_blur_frag.inc": frag(ref, "GTAO/blur.frag"),
        "blurX_frag.inc": frag(ref, "Lighting/blurX.frag"),
        "blurY_frag.inc": frag(ref, "Lighting/blurY.frag"),
        "aggregateLights_frag.inc": frag(ref, "Lighting/aggregateLights.frag"),
        "color_frag.inc": frag(ref, "GTAO/color.frag"),
    }
    files["voxel_gs.inc"], files["voxel_ps.inc"] = voxel_stages(ref)
    for name, text in files.items():
        with open(os.path.join(out, name), "w") as f:
            f.write(text)
    print("make_ref_shaders:", ", ".join(sorted(files)), "->", out)


if __name__ == "__main__":
    main()
