#!/usr/bin/env python3
"""Regenerates include/orbfe_brief_pattern.inc from the reference's BRIEF table (data, not code).

Run only in the build container (needs /root/reference). The table is the learned rBRIEF pattern
at reference src/ORBextractor.cc:197-455; we keep the 1024 integers and drop everything else.
"""
import hashlib, re, sys
ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/src/ORBextractor.cc'
lines = open(ref).read().split('\n')[197:454]
txt = re.sub(r'/\*.*?\*/', '', '\n'.join(lines))
nums = [int(x) for x in re.findall(r'-?\d+', txt)]
assert len(nums) == 1024, len(nums)
h = hashlib.sha256(','.join(map(str, nums)).encode()).hexdigest()
assert h == '88df8ca875cc8db56799edd57bb914edad8acb2d48c202b7a464a575b55dbdb8', h
out = open('include/orbfe_brief_pattern.inc', 'w')
out.write("/* rBRIEF-256 learned sampling pattern: 256 point pairs, 1024 signed offsets in [-13,12],\n"
          "   order x0,y0,x1,y1 per pair. DATA extracted by tools/gen_brief_pattern.py from the table at\n"
          "   reference src/ORBextractor.cc:197-455 (bit_pattern_31_).\n"
          "   sha256 of the comma-joined decimal values: %s */\n" % h)
for i in range(0, 1024, 32):
    out.write(' '.join('%d,' % v for v in nums[i:i + 32]) + '\n')
