"""Decode the scheduling control fields of sm_100a SASS (cuobjdump -sass) and print them beside each instruction:
stall count, yield, write/read scoreboard slot, wait mask.  Used to read the critical path of the latency-bound
mixing kernels without a GPU (see /opt/skills/guides/B300_MICROARCH.md "Single-warp issue model").
Usage: python tools/sass_ctl.py <object-or-so> <function-substring> [start_addr end_addr]"""
import re, subprocess, sys


def decode(path, fn):
    out = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout.splitlines()
    ins = []
    active = False
    pend = None
    for line in out:
        if 'Function :' in line:
            active = fn in line
            continue
        if not active:
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/', line)
        if m:
            pend = (int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16))
            continue
        m = re.match(r'\s*/\* 0x([0-9a-f]{16}) \*/', line)
        if m and pend:
            hi = int(m.group(1), 16)
            ins.append(dict(addr=pend[0], text=pend[1], lo=pend[2], hi=hi, stall=(hi >> 41) & 0xf, yld=(hi >> 45) & 1,
                            wbar=(hi >> 46) & 7, rbar=(hi >> 49) & 7, wait=(hi >> 52) & 0x3f))
            pend = None
    return ins


if __name__ == '__main__':
    ins = decode(sys.argv[1], sys.argv[2])
    lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
    for i in ins:
        if lo <= i['addr'] <= hi:
            w = ''.join(str(b) if (i['wait'] >> b) & 1 else '-' for b in range(6))
            print('%05x  st=%2d %s w%s r%s wait=%s  %s' % (i['addr'], i['stall'], 'Y' if i['yld'] else ' ',
                  i['wbar'] if i['wbar'] < 7 else '-', i['rbar'] if i['rbar'] < 7 else '-', w, i['text']))
