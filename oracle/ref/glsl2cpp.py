#!/usr/bin/env python
"""glsl2cpp.py — turns one of the reference's UNMODIFIED GLSL programs into a C++ translation unit that compiles
against oracle/ref/glsl_compat.h. TEST INFRASTRUCTURE ONLY (see that header).

    python glsl2cpp.py --shader-dir /root/reference/DynamicRadianceVolume/shader --shader cacheGather.comp \
        --define INDDIFFUSE_VIA_SH1 --define ADDRESSVOL_CASCADE_TRANSITIONS --name gather_sh1_t \
        --harness harness_gather.inc --out ../_ref/gen/gather_sh1_t.cpp

The rewrite is purely mechanical (nothing shader-specific lives here); what it does, in order:
  * `#version` is dropped, `#include "x"` is inlined recursively (with #line markers so g++ diagnostics and the
    citations in tests point at the reference file and line); every other preprocessor line — the reference's own
    #define / #ifdef option switches — is left to the C preprocessor;
  * floating literals get an `f` suffix (GLSL literals are single precision);
  * multi-component swizzles `.xyz` / `.rgb` / ... become `.swz<0,1,2>()` (single components are plain members),
    swizzle assignments `v.rgb = e;` become `v.set_swz<0,1,2>(e);`;
  * interface blocks (`layout(...) uniform Name { ... };`, `... buffer Name { ... };`) are dissolved: their members
    become static members of the shader struct, unsized arrays `T[] name;` become `ssbo<T> name;` (bounds-checked);
  * opaque uniforms (`layout(binding=..) uniform sampler2D X;`, images) and `layout(location=..) uniform T x;`
    become static members; `shared T x[..];` becomes a thread-local static (one work group runs on one OS thread);
  * `in T x;` / `out T x;` / `layout(location=..) out T x;` become per-invocation members;
  * `layout(local_size_x = .., ...) in;` becomes three constants; `discard;` sets a flag and returns;
  * parameter qualifiers: `in T p` -> `T p`, `out T p` / `inout T p` -> `T& p`.
The whole text is wrapped in `struct Shader : glsl::Invocation { ... };` inside its own namespace, followed by the
harness (hand-written in this repo: it binds inputs, dispatches and copies results out).
"""
import argparse
import os
import re
import sys

SWZ = {c: i for i, c in enumerate("xyzw")}
SWZ.update({c: i for i, c in enumerate("rgba")})
SWZ.update({c: i for i, c in enumerate("stpq")})

FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")
SWZ_ASSIGN = re.compile(r"\b([A-Za-z_]\w*(?:\[[^\]]*\])?)\.([xyzw]{2,4}|[rgba]{2,4}|[stpq]{2,4})\s*=(?!=)\s*([^;]+);")
SWZ_READ = re.compile(r"\.([xyzw]{2,4}|[rgba]{2,4}|[stpq]{2,4})\b(?!\s*\()")
BLOCK_HEAD = re.compile(r"^\s*layout\s*\([^)]*\)\s*(?:\w+\s+)*(uniform|buffer)\s+(\w+)\s*(\{)?\s*(//.*)?$")
OPAQUE_UNIFORM = re.compile(r"^\s*layout\s*\([^)]*\)\s*(?:(?:restrict|coherent|readonly|writeonly|volatile)\s+)*uniform\s+(\w+)\s+(\w+)\s*;(.*)$")
LOCAL_SIZE = re.compile(r"^\s*layout\s*\(\s*local_size_x\s*=\s*([^,]+),\s*local_size_y\s*=\s*([^,]+),\s*local_size_z\s*=\s*([^)]+)\)\s*in\s*;")
INOUT_DECL = re.compile(r"^\s*(?:layout\s*\([^)]*\)\s*)?(?:flat\s+|smooth\s+|noperspective\s+)?(in|out)\s+(\w+)\s+(\w+)\s*;(.*)$")
SHARED_DECL = re.compile(r"^\s*shared\s+(.*)$")
UNSIZED = re.compile(r"^(\s*)(\w+)\s*\[\s*\]\s+(\w+)\s*;(.*)$")


def swz_indices(letters):
    return ",".join(str(SWZ[c]) for c in letters)


def rewrite_code(line):
    """Rewrites that apply to every non-preprocessor line."""
    line = FLOAT_LIT.sub(lambda m: m.group(1) + "f", line)
    line = SWZ_ASSIGN.sub(lambda m: "%s.set_swz<%s>(%s);" % (m.group(1), swz_indices(m.group(2)), m.group(3).strip()), line)
    line = SWZ_READ.sub(lambda m: ".swz<%s>()" % swz_indices(m.group(1)), line)
    line = re.sub(r"\bdiscard\s*;", "{ gl_Discarded = true; return; }", line)
    # parameter qualifiers (only after '(' or ',')
    line = re.sub(r"([(,]\s*)(?:out|inout)\s+(\w+)\s+(\w+)", r"\1\2& \3", line)
    line = re.sub(r"([(,]\s*)in\s+(\w+)\s+(\w+)", r"\1\2 \3", line)
    return line


def inline_includes(path, shader_dir, out, seen_depth=0):
    if seen_depth > 16:
        raise RuntimeError("#include nesting too deep at " + path)
    name = os.path.basename(path)
    out.append('#line 1 "%s"' % name)
    with open(path, "r", encoding="utf-8", errors="replace") as f:
        lines = f.read().replace("\r\n", "\n").split("\n")
    for no, line in enumerate(lines, 1):
        s = line.strip()
        if s.startswith("#version"):
            out.append("")
            continue
        m = re.match(r'#\s*include\s+"([^"]+)"', s)
        if m:
            inline_includes(os.path.join(shader_dir, m.group(1)), shader_dir, out, seen_depth + 1)
            out.append('#line %d "%s"' % (no + 1, name))
            continue
        out.append(line)


def translate(lines):
    out = []
    in_block = False       # inside a dissolved interface block
    pending_block = False  # saw the block head, waiting for '{'
    in_comment = False
    for line in lines:
        s = line.strip()
        # track /* */ comments coarsely (the shaders only use them around whole statements / directive groups)
        if in_comment:
            out.append(line)
            if "*/" in line:
                in_comment = False
            continue
        if "/*" in line and "*/" not in line.split("/*", 1)[1]:
            in_comment = True
            out.append(line)
            continue
        if s.startswith("#"):
            # preprocessor: keep; numeric macros are single precision too
            if re.match(r"#\s*define\b", s):
                line = FLOAT_LIT.sub(lambda m: m.group(1) + "f", line)
            out.append(line)
            continue
        if pending_block:
            if s.startswith("{"):
                pending_block = False
                in_block = True
                out.append("// {")
                continue
        if in_block:
            if s.startswith("};") or s == "}":
                in_block = False
                out.append("// };")
                continue
            if not s or s.startswith("//"):
                out.append(line)
                continue
            m = UNSIZED.match(line)
            if m:
                out.append("%sinline static ssbo<%s> %s;%s" % (m.group(1), m.group(2), m.group(3), m.group(4)))
            else:
                out.append("inline static " + rewrite_code(line).lstrip())
            continue
        m = LOCAL_SIZE.match(line)
        if m:
            out.append("static constexpr int gl_WorkGroupSize_x = %s, gl_WorkGroupSize_y = %s, gl_WorkGroupSize_z = %s;"
                       % (m.group(1).strip(), m.group(2).strip(), m.group(3).strip()))
            continue
        m = OPAQUE_UNIFORM.match(line)
        if m:
            out.append("inline static %s %s;%s" % (m.group(1), m.group(2), m.group(3)))
            continue
        m = BLOCK_HEAD.match(line)
        if m:
            out.append("// interface block %s %s" % (m.group(1), m.group(2)))
            if m.group(3):
                in_block = True
            else:
                pending_block = True
            continue
        m = INOUT_DECL.match(line)
        if m:
            out.append("%s %s;%s" % (m.group(2), m.group(3), m.group(4)))
            continue
        m = SHARED_DECL.match(line)
        if m:
            out.append("inline static thread_local " + rewrite_code(m.group(1)))
            continue
        out.append(rewrite_code(line))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shader-dir", required=True)
    ap.add_argument("--shader", required=True)
    ap.add_argument("--define", action="append", default=[], help="NAME or NAME=VALUE, as the host passes to the shader compiler")
    ap.add_argument("--name", required=True, help="namespace / symbol suffix of this variant")
    ap.add_argument("--harness", required=True, help="harness .inc appended after the shader struct")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    raw = []
    inline_includes(os.path.join(a.shader_dir, a.shader), a.shader_dir, raw)
    body = translate(raw)
    here = os.path.dirname(os.path.abspath(__file__))
    with open(a.out, "w") as f:
        f.write("// GENERATED by oracle/ref/glsl2cpp.py from the reference's %s — build artefact, never committed.\n" % a.shader)
        f.write('#include "%s"\n' % os.path.join(here, "glsl_compat.h"))
        f.write('#include "%s"\n' % os.path.join(here, "ref_api.h"))
        for d in a.define:
            k, _, v = d.partition("=")
            f.write("#define %s %s\n" % (k, v))
        f.write("#define REF_VARIANT %s\n" % a.name)
        f.write("namespace glsl { namespace %s {\n" % a.name)
        f.write("struct Shader : Invocation {\n")
        f.write("\n".join(body))
        f.write("\n};\n")
        f.write('#line 1 "%s"\n' % a.harness)
        with open(os.path.join(here, a.harness)) as h:
            f.write(h.read())
        f.write("\n} }\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
