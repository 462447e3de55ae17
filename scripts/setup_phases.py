import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import femder_b200 as fd
g=fd.shoebox_tet4(99,133,74)
s=fd.System(0)
for rep in range(3):
    t=time.perf_counter(); s.set_volume_mesh(g.nos,g.elem_vol); s.synchronize(); t1=time.perf_counter()-t
    t=time.perf_counter(); s.assemble_volume(1.21,343.0); s.synchronize(); t2=time.perf_counter()-t
    t=time.perf_counter(); s.set_surface_mesh(g.elem_surf,g.domain_index_surf,np.arange(2,8)); s.synchronize(); t3=time.perf_counter()-t
    t=time.perf_counter(); a=s.heron_areas(); t4=time.perf_counter()-t
    t=time.perf_counter(); s.assemble_surface(); s.synchronize(); t5=time.perf_counter()-t
    print(f"set_volume_mesh {t1:.3f} assemble_volume {t2:.3f} set_surface_mesh {t3:.3f} heron {t4:.3f} assemble_surface {t5:.3f}")
