"""Print the order of barrier tests (W), MMAs (M) and commits (C) in a kernel's SASS."""
import re, subprocess, sys
lib = "/root/repo/super-resolution-building-height-estimation_b200/lib/libbhsr.so"
txt = subprocess.check_output(['cuobjdump', '-sass', lib]).decode()
for f in re.split(r'\n\s*Function : ', txt):
    name = f.split('\n')[0]
    if sys.argv[1] in name:
        lines = [re.sub(r'/\*[0-9a-f]{4,6}\*/\s*', '', l.strip())[:90] for l in f.split('\n') if re.search(r'/\*[0-9a-f]{4,6}\*/', l)]
        seq = ''.join('W' if 'TRYWAIT' in l else 'M' if 'UTCHMMA' in l else 'C' if 'UTCBAR' in l else 'b' if re.match(r'@!?U?P\d+\s+BRA', l) else '' for l in lines)
        print(name[:80]); print(seq)
