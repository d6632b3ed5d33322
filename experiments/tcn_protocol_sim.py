# logical simulation of the tcn kernel's barrier protocol (zero-latency MMAs / TMA): finds protocol deadlocks
import sys
SLOTS, nky, nkx, nsteps, ES, PB = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), 2, 3
class Bar:
    def __init__(s, count): s.count=count; s.pending=count; s.phase=0
    def arrive(s):
        s.pending-=1
        if s.pending==0: s.pending=s.count; s.phase+=1
    def done(s, parity):   # try_wait.parity: true if the phase with this parity has completed = current phase parity != parity
        return (s.phase & 1) != parity
full=[Bar(1) for _ in range(8)]; empty=[Bar(1) for _ in range(8)]
tfull=[Bar(1),Bar(1)]; tempty=[Bar(1),Bar(1)]; hfull=[Bar(1),Bar(1)]; hempty=[Bar(1),Bar(1)]
hready=Bar(1); ufull=[Bar(1),Bar(1)]; pfull=[Bar(1) for _ in range(PB)]; pempty=[Bar(1) for _ in range(PB)]
def wait(b,p,tag):
    while not b.done(p): yield tag
def producer():
    slot=0; phase=0
    def load(n):
        nonlocal slot, phase
        for _ in range(n):
            yield from wait(empty[slot], phase^1, ('prod empty',slot,phase^1))
            full[slot].arrive()
            slot+=1
            if slot==SLOTS: slot=0; phase^=1
    def drain(m):
        yield from wait(pfull[m%PB], (m//PB)&1, ('prod pfull',m))
        pempty[m%PB].arrive()
    if nsteps>0: yield from load(nky)
    if nsteps>1: yield from load(nky)
    for st in range(nsteps):
        yield from load(nkx)
        if st+2<nsteps: yield from load(nky)
        if st>=PB: yield from drain(st-PB)
    for m in range(max(0,nsteps-PB), nsteps): yield from drain(m)
def issuerA():
    slot=0; phase=0
    def skip(n):
        nonlocal slot, phase
        for _ in range(n):
            slot+=1
            if slot==SLOTS: slot=0; phase^=1
    def A(n):
        nonlocal slot, phase
        hb=n&1
        yield from wait(hempty[hb], ((n>>1)&1)^1, ('A hempty',n))
        for _ in range(nky):
            yield from wait(full[slot], phase, ('A full',n,slot,phase))
            empty[slot].arrive()
            slot+=1
            if slot==SLOTS: slot=0; phase^=1
        hfull[hb].arrive()
    if nsteps>0: yield from A(0)
    if nsteps>1: yield from A(1)
    for st in range(nsteps):
        skip(nkx)
        if st+2<nsteps: yield from A(st+2)
def issuerB():
    slot=0; phase=0
    def skip(n):
        nonlocal slot, phase
        for _ in range(n):
            slot+=1
            if slot==SLOTS: slot=0; phase^=1
    def B(n):
        nonlocal slot, phase
        buf=n&1
        yield from wait(tempty[buf], ((n>>1)&1)^1, ('B tempty',n))
        for _ in range(nkx):
            yield from wait(full[slot], phase, ('B full',n,slot,phase))
            empty[slot].arrive()
            slot+=1
            if slot==SLOTS: slot=0; phase^=1
        yield from wait(hready, n&1, ('B hready',n))
        tfull[buf].arrive()
    def P(m):
        pb=m%PB
        yield from wait(pempty[pb], ((m//PB)&1)^1, ('P pempty',m))
        yield from wait(ufull[m&1], (m>>1)&1, ('P ufull',m))
        pfull[pb].arrive()
    skip(min(nsteps,2)*nky)
    for st in range(nsteps):
        yield from B(st)
        if st+2<nsteps: skip(nky)
        if st>=1: yield from P(st-1)
    if nsteps>0: yield from P(nsteps-1)
def epilogue():
    for n in range(nsteps+1):
        do_cv = n<nsteps; do_ep = n>0; m=n-1
        if do_ep: yield from wait(tfull[m&1], (m>>1)&1, ('E tfull',m))
        if do_cv: yield from wait(hfull[n&1], (n>>1)&1, ('E hfull',n))
        if do_cv: hempty[n&1].arrive()
        if do_ep: tempty[m&1].arrive()
        if do_cv: hready.arrive()
        if do_ep: ufull[m&1].arrive()
roles={'prod':producer(),'A':issuerA(),'B':issuerB(),'epi':epilogue()}
state={}
import itertools
for it in range(10**6):
    progressed=False
    for k,g in list(roles.items()):
        try:
            tag=next(g)
            if state.get(k)!=tag: progressed=True
            state[k]=tag
        except StopIteration:
            del roles[k]; progressed=True; state[k]='done'
    if not roles: print('all done'); break
    if not progressed:
        # second pass to confirm
        stuck=True
        for k,g in list(roles.items()):
            tag=next(g, 'done')
            if tag!=state[k]: stuck=False; state[k]=tag
        if stuck: print('DEADLOCK', state); break
