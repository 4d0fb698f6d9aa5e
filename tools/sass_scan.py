"""Scans SASS for the ptxas waterfall-loop hazard: inside a non-uniform-texture-handle loop (R2UR ... BRA.U.ANY) an
unpredicated write to the register that is also the TEX destination clobbers lanes served by an earlier iteration."""
import re, sys, subprocess
def scan(path, name_filter=None):
    out = subprocess.run(["cuobjdump","-sass",path],capture_output=True,text=True).stdout
    funcs = re.split(r'\n\s*Function : ', out)
    bad = []
    for f in funcs[1:]:
        fname = f.split('\n',1)[0].strip()
        lines=[l for l in f.split('\n') if re.search(r'/\*[0-9a-f]{4}\*/',l)]
        addr=lambda x:int(re.search(r'/\*([0-9a-f]{4})\*/',x).group(1),16)
        for i,l in enumerate(lines):
            m=re.search(r'BRA\.U\.ANY (0x[0-9a-f]+)',l)
            if not m: continue
            tgt=int(m.group(1),16)
            body=[x for x in lines[:i+1] if addr(x)>=tgt]
            for t in [x for x in body if re.search(r'\bTEX|\bTLD',x)]:
                dest=re.search(r'T(?:EX|LD)\S* \S+, (R\d+),',t)
                if not dest: continue
                d=dest.group(1); n=int(d[1:])
                # dest may be a vector (up to 4 regs); check scalar + following regs conservatively
                for x in body:
                    if x is t: continue
                    w=re.search(r'^\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?(\S+)\s+(R\d+)',x)
                    if w and w.group(2)==d and not w.group(1).startswith(('TEX','TLD','ST','BRA','R2UR')) :
                        pred = re.search(r'/\*[0-9a-f]{4}\*/\s+(@!?P\d+)',x)
                        if not pred: bad.append((fname, hex(tgt), d, x.strip()[:70]))
    return bad
if __name__ == '__main__':
    for p in sys.argv[1:]:
        b=scan(p); print(p, 'suspicious waterfall clobbers:', len(b))
        for x in b[:12]: print('   ',x)
