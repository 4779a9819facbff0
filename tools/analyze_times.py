import sys, numpy as np
launches=[]; cur=None
for line in open(sys.argv[1]):
    if line.startswith('launch'):
        cur=[line.strip(),[]]; launches.append(cur)
    else:
        cur[1].append([int(x) for x in line.split()])
for name,rows in launches[:int(sys.argv[2]) if len(sys.argv)>2 else 40]:
    a=np.array(rows,dtype=np.int64)
    ok=a[:,0]>0
    a=a[ok]
    t0=a[:,0].min()
    dur=(a[:,6].max()-t0)/1e3
    pro=(a[:,1]-a[:,0]).mean()/1e3; first=(a[:,2]-a[:,1]).mean()/1e3; main=(a[:,3]-a[:,2]).mean()/1e3
    drain=(a[:,4]-a[:,3]).mean()/1e3; epi=(a[:,5]-a[:,4]).mean()/1e3; tail=(a[:,6]-a[:,5]).mean()/1e3
    life=(a[:,6]-a[:,0]).mean()/1e3
    stg=((a[:,7]-a[:,4]).mean()/1e3) if a.shape[1]>7 and (a[:,7]>0).all() else float('nan')
    print(f"{name}: kernel {dur:7.1f} us | per-CTA us: prologue {pro:5.2f} firstTMA {first:5.2f} mainloop(issue) {main:6.2f} mma-drain {drain:5.2f} epilogue {epi:5.2f} (tmem->smem {stg:5.2f}) teardown {tail:5.2f} | life {life:6.2f} ctas {len(a)}")
