"""Turn a source file into a C++ translation unit defining `extern const char NAME[]` (used to hand
mfit_eval.cuh to NVRTC at run time)."""
import sys

path, name = sys.argv[1], sys.argv[2]
text = open(path).read()
print("// generated from %s -- do not edit" % path)
print('extern "C++" { extern const char %s[]; }' % name)
print("const char %s[] =" % name)
for line in text.split("\n"):
    esc = line.replace("\\", "\\\\").replace('"', '\\"')
    print('  "%s\\n"' % esc)
print(";")
