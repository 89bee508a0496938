#!/usr/bin/env python3
"""embed.py OUT.c FILE... -- turn the reference's .cl files into byte arrays of a C file that is
written under oracle/_ref/ (git-ignored build output)."""
import os
import sys

out, files = sys.argv[1], sys.argv[2:]
with open(out, "w") as f:
    f.write("/* GENERATED build artefact: reference OpenCL sources embedded for oracle/_ref/mipref_ocl */\n")
    f.write("#include <stddef.h>\n")
    for i, p in enumerate(files):
        data = open(p, "rb").read()
        f.write(f"static const unsigned char src{i}[] = {{")
        f.write(",".join(str(b) for b in data))
        f.write("};\n")
    f.write(f"const int ref_cl_count = {len(files)};\n")
    f.write("const char* const ref_cl_names[] = {" + ",".join(f'"{os.path.basename(p)}"' for p in files) + "};\n")
    f.write("const unsigned char* const ref_cl_data[] = {" + ",".join(f"src{i}" for i in range(len(files))) + "};\n")
    f.write("const size_t ref_cl_size[] = {" + ",".join(f"sizeof(src{i})" for i in range(len(files))) + "};\n")
