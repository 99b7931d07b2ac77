"""TEST INFRASTRUCTURE ONLY -- host build of the CUDA sources for the SIMT interpreter (tests/simt/shim/simt.h).

`build()` translates multical_b200/csrc/*.cu[h] textually (kernel launches -> simt::launch, `__shared__` -> per-block storage,
the three inline-PTX statements -> their C++ meaning), compiles the result with g++ against the stand-in cuda_runtime.h and
returns the path of tests/simt/build/libmcba_simt.so, which exports the same C-ABI as multical_b200/libmcba.so.  The product
never loads this library (multical_b200/_native.py only knows libmcba.so); tests/test_simt_kernels.py points the ctypes binding
at it so that the CPU suite (-m "not gpu") runs the real kernels and the real host driver at small sizes.
"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "multical_b200", "csrc")
OUT = os.path.join(HERE, "build")
LIB = os.path.join(OUT, "libmcba_simt.so")


def _match_back(s, i, open_c, close_c):
  """index of the `open_c` matching the `close_c` at s[i]."""
  depth = 0
  while i >= 0:
    if s[i] == close_c: depth += 1
    elif s[i] == open_c:
      depth -= 1
      if depth == 0: return i
    i -= 1
  raise ValueError("unbalanced")


def _match_fwd(s, i, open_c, close_c):
  depth = 0
  while i < len(s):
    if s[i] == open_c: depth += 1
    elif s[i] == close_c:
      depth -= 1
      if depth == 0: return i
    i += 1
  raise ValueError("unbalanced")


def _split_top(s):
  parts, depth, cur = [], 0, ""
  for ch in s:
    if ch in "([{": depth += 1
    elif ch in ")]}": depth -= 1
    if ch == "," and depth == 0: parts.append(cur.strip()); cur = ""
    else: cur += ch
  parts.append(cur.strip())
  return parts


def rewrite_launches(src):
  out, pos = "", 0
  while True:
    i = src.find("<<<", pos)
    if i < 0: return out + src[pos:]
    # kernel expression: identifier, optionally followed by a balanced <...> template argument list
    j = i - 1
    while src[j].isspace(): j -= 1
    if src[j] == ">": j = _match_back(src, j, "<", ">") - 1
    while j >= 0 and (src[j].isalnum() or src[j] in "_:"): j -= 1
    kernel = src[j + 1:i].strip()
    k = src.find(">>>", i)
    cfg = _split_top(src[i + 3:k])
    a0 = src.find("(", k)
    assert src[k + 3:a0].strip() == "", f"unexpected text after launch configuration of {kernel}"
    a1 = _match_fwd(src, a0, "(", ")")
    args = src[a0 + 1:a1]
    grid, block = cfg[0], cfg[1]
    smem = cfg[2] if len(cfg) > 2 else "0"
    name = kernel.replace('"', "")
    out += src[pos:j + 1] + f'simt::launch("{name}", dim3({grid}), dim3({block}), (size_t)({smem}), [&]() {{ {kernel}({args}); }})'
    pos = a1 + 1


ASM = re.compile(r"asm\s+volatile\s*\((.*?)\)\s*;", re.S)


def rewrite_asm(m):
  body = m.group(1)
  if "mma.sync.aligned.m8n8k4" in body: return "simt::dmma884(c0, c1, a, b);"
  # a volatile load is how the kernels poll a flag written by another rank: let the other threads of the block run between polls (on
  # hardware every lane makes progress; a fiber that spins without yielding would starve the lanes that publish this rank's own flags)
  if "ld.volatile.global.u64" in body: return "v = *(volatile const unsigned long long*)p; simt::spin_yield();"
  if "st.volatile.global.u64" in body: return "*(volatile unsigned long long*)p = v;"
  if "globaltimer" in body: return "t_ = 0;"                       # profiling timestamps: no clock on the interpreter
  if "red.global.add.f64" in body: return "*p += v;"
  # bulk asynchronous copies + mbarrier (csrc/linearize.cuh): the interpreter copies synchronously, the barrier has nothing left to wait for
  # bulk asynchronous copies + mbarrier (csrc/solver_kernels.cuh): the copy is synchronous here, so a phase completes when the producer
  # arms it (the word counts completed phases); a consumer waiting for parity P yields until the count's parity differs from P
  if "cp.async.bulk.shared" in body: return "memcpy(dst, src, bytes);"
  if "mbarrier.init" in body: return "*bar = 0ull;"
  if "mbarrier.arrive.expect_tx" in body: return "*(volatile unsigned long long*)bar = *bar + 1ull;"
  if "mbarrier.try_wait" in body: return "while ((*(volatile unsigned long long*)bar & 1ull) == (unsigned long long)phase) simt::spin_yield();"
  if "fence.mbarrier_init" in body or "fence.proxy.async" in body: return ";"                # fire-and-forget fp64 add to an address only this thread updates
  raise ValueError("inline PTX without a host meaning: " + body[:80])


def translate(src):
  src = rewrite_launches(src)
  src = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\s*\[\s*\]\s*;", r"\1* \2 = (\1*)simt::dyn_smem();", src)
  src = re.sub(r"\b__shared__\b", "static thread_local", src)       # per block, blocks of one rank run one after the other; ranks are host threads
  src = ASM.sub(rewrite_asm, src)
  src = src.replace('#include "../../include/mcba.h"', f'#include "{os.path.join(ROOT, "include", "mcba.h")}"')
  return src


def sources():
  return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def stale():
  if not os.path.exists(LIB): return True
  t = os.path.getmtime(LIB)
  deps = [os.path.join(CSRC, f) for f in sources()] + [os.path.join(ROOT, "include", "mcba.h"), os.path.abspath(__file__)]
  shim = os.path.join(HERE, "shim")
  for d, _, fs in os.walk(shim): deps += [os.path.join(d, f) for f in fs]
  return any(os.path.getmtime(d) > t for d in deps)


def build(force=False):
  if not (force or stale()): return LIB
  os.makedirs(OUT, exist_ok=True)
  main = None
  for f in sources():
    with open(os.path.join(CSRC, f)) as fh: text = translate(fh.read())
    dst = os.path.join(OUT, f[:-3] + ".cpp" if f.endswith(".cu") else f)
    with open(dst, "w") as fh: fh.write(text)
    if f.endswith(".cu"):
      main = dst
      # marks the library as the interpreter build: multical_b200/_native.py refuses to load it unless a test asked for it
      with open(dst, "a") as fh: fh.write('\nextern "C" int mcba_simt_build(void) { return 1; }\n')
  tmp = LIB + f".{os.getpid()}.tmp"            # never overwrite a library that a running test process may have mapped
  cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-fno-strict-aliasing", "-w", "-Wno-unknown-pragmas",
         "-I", os.path.join(HERE, "shim"), "-o", tmp, main, "-ldl"]
  subprocess.run(cmd, check=True, cwd=OUT)
  os.replace(tmp, LIB)
  return LIB


if __name__ == "__main__":
  import sys
  print(build(force="--force" in sys.argv))
