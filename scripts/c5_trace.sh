python - <<'P'
import numpy as np, os
rng=np.random.default_rng(1)
g=np.frombuffer(b"ACGT",dtype=np.uint8)[rng.integers(0,4,size=64_000_000,dtype=np.uint8)]
tot=0;n=0
with open('/tmp/t.fq','wb',buffering=1<<24) as f:
    while tot<2_000_000_000:
        L=max(200,int(rng.lognormal(0,0.6)*100000/np.exp(0.36))); a=int(rng.integers(0,len(g)-L)); s=g[a:a+L]; n+=1
        f.write(b"@r%d\n"%n+s.tobytes()+b"\n+\n"+b"I"*L+b"\n"); tot+=L
print('reads',n,'GB',os.path.getsize('/tmp/t.fq')/1e9)
P
cat /tmp/t.fq > /dev/null
for g in 1 2; do echo "== GPUS=$g"; ( time CORNETTO_GPUS=$g CORNETTO_TRACE=1 cornetto_b200/bin/cornetto telofind /tmp/t.fq > /tmp/t.out ) 2>&1 | grep -v "^$" | tail -40; done
echo "== GPUS=1 serial reader"; ( time CORNETTO_INGEST=0 cornetto_b200/bin/cornetto telofind /tmp/t.fq > /tmp/t.out2 ) 2>&1 | tail -4; cmp /tmp/t.out /tmp/t.out2 && echo same
