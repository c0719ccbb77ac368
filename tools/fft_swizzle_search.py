#!/usr/bin/env python3
"""Developer tool: bank-conflict model of the 512-point FFT buffer of t41rx_stream_rx_kernel; exhaustive search of
padding / XOR-swizzle index maps (result: i ^ ((i >> 3) & 15), 304 wavefronts = the minimum)."""
import itertools
def wf(addrs_words, width_words):
    # addrs: list of 32 lane word addresses for an access of width_words (1,2,4); returns wavefronts
    # model: hardware splits into groups so that each group moves <=128B: for 64-bit: 2 half-warps; 128-bit: 4 quarter-warps
    groups = {1:[range(32)], 2:[range(0,16), range(16,32)], 4:[range(0,8),range(8,16),range(16,24),range(24,32)]}[width_words]
    total=0
    for g in groups:
        banks={}
        for l in g:
            a=addrs_words[l]
            for w in range(width_words):
                b=(a+w)%32
                banks.setdefault(b,set()).add(a+w)
        total+=max(len(v) for v in banks.values())
    return total
def fft_cost(pos):
    tot=0
    for warp in range(2):
        bs=[32*warp+l for l in range(32)]
        # pass A: i0=b, n2=64: loads/stores element b+64m
        for m in range(8): tot+=2*wf([2*pos(b+64*m) for b in bs],2)
        # pass B: i0=(b>>3)*64+(b&7), n2=8
        for m in range(8): tot+=2*wf([2*pos((b>>3)*64+(b&7)+8*m) for b in bs],2)
        # mid: 8b+m
        for m in range(8): tot+=2*wf([2*pos(8*b+m) for b in bs],2)
        # inv B, inv A (stores only upper half for A')
        for m in range(8): tot+=2*wf([2*pos((b>>3)*64+(b&7)+8*m) for b in bs],2)
        for m in range(8): tot+=wf([2*pos(b+64*m) for b in bs],2)
        for m in range(4,8): tot+=wf([2*pos(b+64*m) for b in bs],2)
    return tot
cands={
 'i+(i>>3)': lambda i: i+(i>>3),
 'i+(i>>4)': lambda i: i+(i>>4),
 'i+(i>>3)+(i>>6)': lambda i: i+(i>>3)+(i>>6),
 'i+(i>>4)+(i>>6)': lambda i: i+(i>>4)+(i>>6),
 'i+(i>>3)+(i>>7)': lambda i: i+(i>>3)+(i>>7),
 'i+2*(i>>4)': lambda i: i+2*(i>>4),
 'i+(i>>3)+2*(i>>6)': lambda i: i+(i>>3)+2*(i>>6),
 'i+(i>>3)+(i>>6)+(i>>9)': lambda i: i+(i>>3)+(i>>6),
 'i': lambda i:i,
 'i+(i>>4)+(i>>7)': lambda i: i+(i>>4)+(i>>7),
 'i+(i>>2)': lambda i: i+(i>>2),
 'i+(i>>5)': lambda i: i+(i>>5),
 'i+(i>>3)-(i>>6)': lambda i: i+(i>>3)-(i>>6),
}
ideal = 2*(8*2*2*4 + 8*2 + 4*2)   # rough
for k,f in cands.items():
    print('%-28s cost %d  size %d' % (k, fft_cost(f), max(f(i) for i in range(512))+1))
print('--- search')
best=[]
for a,b,c,d in itertools.product(range(0,3),range(0,3),range(0,3),range(0,3)):
    f=lambda i,a=a,b=b,c=c,d=d: i+a*(i>>3)+b*(i>>4)+c*(i>>5)+d*(i>>6)
    size=max(f(i) for i in range(512))+1
    if size>600: continue
    best.append((fft_cost(f),size,(a,b,c,d)))
best.sort()
print(best[:8])
# xor swizzles on float2 index: i ^ ((i>>s)&m)
bx=[]
for s in range(3,9):
  for m in (1,3,7,15):
    for s2 in (0,3,4,5,6,7):
      for m2 in (0,1,3,7,15):
        f=lambda i,s=s,m=m,s2=s2,m2=m2: i ^ ((i>>s)&m) ^ (((i>>s2)&m2) if m2 else 0)
        if len(set(f(i) for i in range(512)))!=512: continue
        bx.append((fft_cost(f),(s,m,s2,m2)))
bx.sort()
print(bx[:8])
print('--- other patterns for xor swizzle vs current')
def others(pos):
    tot=0
    for o in range(8):   # assembly: scalar word writes, lane L element 8L+o and 256+8L+o (component w2=0)
        tot+=wf([2*pos(8*l+o) for l in range(32)],1)
        tot+=wf([2*pos(256+8*l+o) for l in range(32)],1)
    zr=0
    for warp in range(2):
        for o in range(4):
            zr+=wf([2*pos(256+32*warp+l+64*o) for l in range(32)],2)
    return tot,zr
print('current', others(lambda i:i+(i>>3)))
print('xor', others(lambda i:i^((i>>3)&15)))
