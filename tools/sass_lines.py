#!/usr/bin/env python3
"""SASS of one kernel restricted to a range of CUDA source lines (companion of tools/ncu_lines.py: the profile names the
hot lines, this shows what the compiler made of them).

    cuobjdump -xelf all odr_audioenc_b200/libtoolame_b200.so        # -> mp2_kernels.sm_100a.cubin
    nvdisasm -g -c mp2_kernels.sm_100a.cubin > all.sass
    tools/sass_lines.py all.sass 'k_filterbankILi2ELi32ELb1' 314 320

KERNEL_REGEX is matched against the mangled section name (a template instance needs its arguments, as above)."""
import re, sys
# usage: sass_lines.py SASSFILE KERNEL_REGEX L0 L1  -> instructions attributed to source lines [L0, L1], in address order
f, kern, l0, l1 = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
inside = False; cur = None; n = 0
for ln in open(f):
    if ln.startswith("//---") and ".text." in ln:
        inside = re.search(kern, ln) is not None
        continue
    if not inside: continue
    m = re.search(r'line (\d+)', ln) if "//##" in ln else None
    if m: cur = int(m.group(1)); continue
    if re.match(r"\s+/\*[0-9a-f]+\*/", ln) and cur is not None and l0 <= cur <= l1:
        print("%4d %s" % (cur, ln.rstrip()[8:90])); n += 1
    elif ln.startswith(".L_") and cur is not None: print("     " + ln.rstrip())
print(n, "instructions")
