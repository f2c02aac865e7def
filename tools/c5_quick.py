import json, os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import bench_c5
r = bench_c5.run_ours(512, 200, 500)
print(os.environ.get('MEDGP_GRAPHS','default'), json.dumps(r['w_update']), json.dumps(r['wo_update']))
