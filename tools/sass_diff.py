#!/usr/bin/env python
"""Compare the SASS of two builds of libcuml_b200.so kernel by kernel.

    python tools/sass_diff.py OLD.so NEW.so

Used to check that a change which only adds opt-in instantiations leaves the kernels that were measured on
hardware untouched.  For every kernel (demangled name) of OLD it reports: identical / same multiset of
instructions (order differs) / different (instruction-count delta), and lists kernels only present on one side.
Addresses, the control-code words and register-allocation-neutral whitespace are stripped before comparing.
"""
import collections
import re
import subprocess
import sys


def kernels(path):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);\s*/\*", line)
        if m and name is not None:
            out[name].append(re.sub(r"\s+", " ", m.group(1)))
    names = list(out)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return {d: out[n] for n, d in zip(names, dem)}


def main():
    old, new = kernels(sys.argv[1]), kernels(sys.argv[2])
    same = reordered = changed = renamed = 0
    for name, ins in old.items():
        short = name if len(name) < 150 else name[:147] + "..."
        if name not in new:
            # a template parameter list that grew renames the instantiation: look for the same instruction stream
            # under a name only the new build has
            twin = next((m for m, o in new.items() if m not in old and o == ins), None) or next(
                (m for m, o in new.items() if m not in old and collections.Counter(o) == collections.Counter(ins)), None)
            if twin:
                renamed += 1
                print("RENAMED   ", len(ins), short, "->", twin[:100], "(identical)" if new[twin] == ins else "(reordered)")
            else:
                print("ONLY OLD  ", short)
            continue
        other = new[name]
        if ins == other:
            same += 1
        elif collections.Counter(ins) == collections.Counter(other):
            reordered += 1
            print("REORDERED ", len(ins), short)
        else:
            changed += 1
            print("CHANGED   ", len(ins), "->", len(other), short)
    for name in new:
        if name not in old:
            print("ONLY NEW  ", len(new[name]), name if len(name) < 150 else name[:147] + "...")
    print(f"identical {same}, reordered {reordered}, renamed {renamed}, changed {changed}, old {len(old)}, new {len(new)}")


if __name__ == "__main__":
    main()
