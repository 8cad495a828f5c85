import sys
sys.path.insert(0, "/root/repo")
from cpu_ray_tracer_b200 import api, host_build
tris = host_build.terrain_mesh(int(sys.argv[1]), seed=1)
print(api.build_bvh_gpu(tris)[2])
