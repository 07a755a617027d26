"""Turn a source file (with its local #include "..." files inlined) into a C++ translation unit defining
`extern const char NAME[]` -- used to hand mfit_eval.cuh to NVRTC at run time."""
import os
import re
import sys

path, name = sys.argv[1], sys.argv[2]
base = os.path.dirname(os.path.abspath(path))


def inline(p, seen):
    out = []
    for line in open(p).read().split("\n"):
        m = re.match(r'\s*#include\s+"([^"]+)"', line)
        if m and os.path.exists(os.path.join(base, m.group(1))):
            q = os.path.join(base, m.group(1))
            if q not in seen:
                seen.add(q)
                out += inline(q, seen)
        else:
            out.append(line)
    return out


text = "\n".join(inline(path, {os.path.abspath(path)}))
print("// generated from %s -- do not edit" % path)
print('extern "C++" { extern const char %s[]; }' % name)
print("const char %s[] =" % name)
for line in text.split("\n"):
    esc = line.replace("\\", "\\\\").replace('"', '\\"')
    print('  "%s\\n"' % esc)
print(";")
