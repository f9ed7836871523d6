"""SASS opcode histogram of one kernel of the built library (cuobjdump, no GPU needed).

    python tools/sass_hist.py gae_ppo_ws_kernelILi6ELb1 [object or .so] > profiles/rNN_xxx_sass.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    pat = sys.argv[1]
    obj = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 'di-engine_b200', 'lib', 'libb200rl.so')
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    hist, name, total = collections.Counter(), None, 0
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = m.group(1)
            continue
        if name is None or pat not in name:
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)', line)
        if m:
            hist[m.group(1)] += 1
            total += 1
    print('# SASS opcode histogram of *%s* in %s (%d instructions)' % (pat, os.path.relpath(obj, ROOT), total))
    for op, n in hist.most_common():
        print('%6d  %s' % (n, op))


if __name__ == '__main__':
    main()
