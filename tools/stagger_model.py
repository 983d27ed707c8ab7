"""Fluid model of one csp timestep under staggered dispatch of the collision class (CPU, numpy):
streamer rate as a function of the collider warps resident beside them, fitted to the warp
timelines of profiles/r02/warp_trace_r2b_csp_step*.txt. See profiles/r02/experiments/README.md:
the GPU did not follow it."""
import numpy as np
# (processed, facets, collisions, hist_ms measured, floor_ms)
steps={3:(1000000,689067338,18605547,3.813,3.503),4:(980008,646861044,51349413,4.633,3.289),5:(924839,640954109,7362207,3.225,3.258),
6:(916929,599916897,52670296,4.520,3.050),7:(860323,573683135,33424506,3.768,2.916),8:(824401,554472048,22943980,3.284,2.819),
9:(799739,551240607,3853415,2.759,2.802),10:(795597,532526241,26170795,3.359,2.707)}
TC=2.3
def r(c):
    return min(2.18, 2.7-0.168*c, (24-c)*0.095)/2.18
def sim(n,floor,ncoll_particles,f,g,dt=0.002):
    c_tot=ncoll_particles/32/148
    n_stream_warps=(n-ncoll_particles)/32/148  # per SM
    cA=c_tot*(1-g); cB=c_tot*g
    t=0.0; done=0.0; startB=None
    while True:
        cres=(cA if t<TC else 0.0)+(cB if (startB is not None and t<startB+TC) else 0.0)
        if startB is None and g>0:
            disp=(done/floor)+ (24-cres)/n_stream_warps
            if disp>=f: startB=t; continue
        if done<floor:
            done+=r(cres)*dt
        t+=dt
        endA=TC; endB=(startB+TC) if startB is not None else 0
        if done>=floor and t>=endA and (g==0 or (startB is not None and t>=endB)): return t,startB
        if t>20: return t,startB
for k,(n,fa,co,ms,floor) in steps.items():
    ncp=co/930.0
    base,_=sim(n,floor,ncp,0,0)
    out=[f"step {k}: share {ncp/n*1000:.0f}permille c={ncp/32/148:.1f} measured {ms:.2f} model {base:.2f} |"]
    for f in (0.2,0.3,0.4,0.5):
        t,sb=sim(n,floor,ncp,f,0.5)
        out.append(f"f={f}: {t:.2f} (B@{sb:.2f})")
    print(" ".join(out))
print()
for k in (4,6,7,10):
    n,fa,co,ms,floor=steps[k]; ncp=co/930.0
    for g in (0.4,0.5,0.6):
        out=[f"step {k} g={g}:"]
        for f in (0.4,0.5,0.55,0.6,0.65,0.7):
            t,sb=sim(n,floor,ncp,f,g)
            out.append(f"f={f}: {t:.2f}")
        print(" ".join(out))
