"""Regenerates the `extern "C"` block of rust/genfer-taylor-sys/src/lib.rs from include/genfer_taylor.h, so the shim
always declares EVERY entry point of the C ABI (tests/test_abi.py::test_rust_shim_declares_every_symbol checks it).
usage: python tools/gen_rust_externs.py [--check]"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "genfer_taylor.h")
LIB = os.path.join(ROOT, "rust", "genfer-taylor-sys", "src", "lib.rs")
BEGIN, END = "    // ---- BEGIN GENERATED (tools/gen_rust_externs.py) ----\n", "    // ---- END GENERATED ----\n"

CTYPES = {
    "int": "c_int", "void": "c_void", "double": "f64", "uint64_t": "u64", "int64_t": "i64", "uint32_t": "u32",
    "size_t": "usize", "char": "c_char", "unsigned": "c_uint",
}


def rust_type(ctype: str) -> str:
    t = ctype.strip()
    const = "const " in (" " + t + " ")
    t = re.sub(r"\bconst\b", "", t).strip()
    stars = t.count("*")
    base = t.replace("*", "").strip()
    if base.startswith("struct "):
        base = base[7:]
    r = CTYPES.get(base, base)
    if stars == 0:
        return r
    # `const T*` -> *const T ; `T**` -> *mut *mut T ; `const T**` -> *mut *const T
    out = r
    for i in range(stars):
        out = ("*const " if (const and i == 0) else "*mut ") + out
    return out


def declarations():
    text = open(HDR).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"#.*", "", text)
    decls = []
    for m in re.finditer(r"([A-Za-z_][\w \*]*?)\b(gt[piu]_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for i, a in enumerate(args.split(",")):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)([A-Za-z_]\w*)$", a)
                ty, nm = (mm.group(1), mm.group(2)) if mm and mm.group(1).strip() else (a, f"a{i}")
                if nm in ("type", "in", "ref", "box", "move", "fn", "mod"):
                    nm += "_"
                params.append(f"{nm}: {rust_type(ty)}")
        r = "" if ret == "void" else f" -> {rust_type(ret)}"
        decls.append((name, f"    pub fn {name}({', '.join(params)}){r};\n"))
    return decls


def main():
    decls = declarations()
    src = open(LIB).read()
    i, j = src.index(BEGIN) + len(BEGIN), src.index(END)
    new = src[:i] + "".join(d for _, d in decls) + src[j:]
    if "--check" in sys.argv:
        sys.exit(0 if new == src else 1)
    open(LIB, "w").write(new)
    print(f"{len(decls)} declarations written")


if __name__ == "__main__":
    main()
