import sys; sys.path.insert(0,".")
import jubjub_b200 as jj, numpy as np, time
from bench import generator_mont, SEED0, timed
from oracle import binding as ob
eng=jj.Engine(0)
gen=generator_mont(eng)
n=1<<20
k=eng.fe_to_bytes("fr", eng.fe_stream("fr", SEED0+2, n, device=True))
o=eng.empty((n,20))
t0=time.perf_counter(); eng.scalar_mul_fixed_vartime(gen,k,out=o); eng.sync(); t1=time.perf_counter()
eng.scalar_mul_fixed_vartime(gen,k,out=o)
ms=timed(eng, lambda: eng.scalar_mul_fixed_vartime(gen,k,out=o), 5)
kh=k.download()[:2000]
ok=(eng.batch_normalize(o.download()[:2000])==ob.batch_normalize(ob.scalar_mul_fixed(ob.generator(),kh))).all()
print(f"{sys.argv[1]}: {ms:.3f} ms  {n/ms*1e3:.3e}/s  first call (table build) {1e3*(t1-t0):.1f} ms  parity {ok}")
