#!/usr/bin/env python
"""Rewrites CUDA kernel launches  kernel<T...><<<grid, block, smem, stream>>>(args)  into  gb_mock::launch(grid, block, [&]() { kernel<T...>(args); })
so that a host compiler can build the file against tests/mock/shim/cuda_runtime.h.  Test infrastructure only.
usage: transform.py in.cu out.cpp"""
import re
import sys


def match_back(text, i, open_c, close_c):
    """text[i] == close_c: index of the matching open_c scanning backwards"""
    depth = 0
    while i >= 0:
        if text[i] == close_c:
            depth += 1
        elif text[i] == open_c:
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced")


def match_fwd(text, i, open_c, close_c):
    depth = 0
    while i < len(text):
        if text[i] == open_c:
            depth += 1
        elif text[i] == close_c:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced")


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


# functions of the product that are nothing but a PTX instruction: their bodies become calls of the emulation in shim/mock_simt.h
ASM_WRAPPERS = ["pk", "upk", "fma2", "mul2", "add2", "sub2", "mbar_init", "mbar_expect_tx", "bulk_g2s", "bulk_g2s_hint", "mbar_wait", "cp_async16", "cp_async_arrive", "mbar_arrive"]
# kernels whose threads cooperate (shared memory, barriers, shuffles): run block by block on fibres (shim/mock_simt.h)
COOPERATIVE = ["dhop_fast_kernel", "dhop_col_kernel", "dhop_col2_kernel", "dhop_col2_kernel_fn", "smat_kernel", "smat_reduce_kernel", "stri_kernel", "pack_send_kernel"]
# kernels that only sometimes need it: C++ condition on the (single) kernel argument.  The generic hopping kernel makes threads 0..7
# acquire the neighbours' epoch flags and everybody else wait at a barrier -- only when it consumes peer-to-peer halos.
COOPERATIVE_IF = {"dhop_kernel": "({arg}).flags != nullptr"}


def replace_asm_wrappers(text):
    for name in ASM_WRAPPERS:
        for m in list(re.finditer(r"__device__\s+__forceinline__\s+[\w:]+\s+" + name + r"\s*\(", text))[::-1]:
            p0 = m.end() - 1
            p1 = match_fwd(text, p0, "(", ")")
            b0 = text.index("{", p1)
            b1 = match_fwd(text, b0, "{", "}")
            if "asm" not in text[b0:b1]:
                continue
            names = [re.findall(r"\w+", a)[-1] for a in split_top(text[p0 + 1:p1])]
            text = text[:b0] + "{ return gb_mock::" + name + "(" + ", ".join(names) + "); }" + text[b1 + 1:]
    return text


def strip_asm(text):
    """the inline PTX left after replace_asm_wrappers: the system-scope flag store / load of the peer-to-peer halos become GCC atomics,
    anything else is dropped"""
    for kw in ("asm volatile(", "asm("):
        while True:
            k = text.find(kw)
            if k < 0:
                break
            e = match_fwd(text, text.index("(", k), "(", ")")
            body = text[k:e + 1]
            ops = re.findall(r'"=?[lrf]"\s*\(', body)
            repl = "(void)0"
            if "st.release.sys.global.u64" in body or "ld.acquire.sys.global.u64" in body:
                args = []
                for mm in re.finditer(r'"=?l"\s*\(', body):
                    a0 = mm.end() - 1
                    args.append(body[a0 + 1:match_fwd(body, a0, "(", ")")])
                if "st.release" in body:
                    repl = f"__atomic_store_n((unsigned long long *)({args[0]}), (unsigned long long)({args[1]}), __ATOMIC_RELEASE)"
                else:
                    repl = f"{args[0]} = __atomic_load_n((const unsigned long long *)({args[1]}), __ATOMIC_ACQUIRE)"
            text = text[:k] + repl + text[e + 1:]
    return text


def shared_memory(text):
    text = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?unsigned char (\w+)\[\];", r"unsigned char *\1 = gb_mock::dynamic_smem();", text)
    return re.sub(r"\b__shared__\b", "static thread_local", text)


def transform(text):
    text = shared_memory(strip_asm(replace_asm_wrappers(text)))
    while True:
        k = text.find("<<<")
        if k < 0:
            return text
        # kernel expression: identifier, optionally followed by <template args>, right before <<<
        j = k - 1
        if text[j] == ">":
            j = match_back(text, j, "<", ">") - 1
        while j >= 0 and (text[j].isalnum() or text[j] in "_:"):
            j -= 1
        start = j + 1
        kernel = text[start:k]
        e = text.index(">>>", k)
        cfg = split_top(text[k + 3:e])
        a0 = text.index("(", e)
        a1 = match_fwd(text, a0, "(", ")")
        args = text[a0:a1 + 1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        base = kernel.split("<")[0].split("::")[-1]
        coop = "true" if base in COOPERATIVE else COOPERATIVE_IF[base].format(arg=args[1:-1]) if base in COOPERATIVE_IF else "false"
        text = text[:start] + f"gb_mock::launch(dim3({cfg[0]}), dim3({cfg[1]}), {smem}, {coop}, \"{kernel}\", [&]() {{ {kernel}{args}; }})" + text[a1 + 1:]


if __name__ == "__main__":
    open(sys.argv[2], "w").write("// generated by tests/mock/transform.py from " + sys.argv[1] + "\n" + transform(open(sys.argv[1]).read()))
