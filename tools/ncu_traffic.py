"""profiles/r02_ncu_traffic.json from the `ncu --set full` captures of tools/capture_profiles.sh (run where ncu is
installed; no GPU needed): dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels bench.py's roofline
names.  er_embedding_bwd is several launches (memset, count, place, fused, big, long); the full capture holds the two
that move DRAM bytes that matter (place: the pairs, fused: gradient rows + table rows) - the sum of their averages is
recorded, and the text summaries under profiles/ list each.

python tools/ncu_traffic.py gpurun_out profiles/r02_ncu_traffic.json"""
import csv
import json
import os
import subprocess
import sys


def launches(path):
  out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units = rows[0], rows[1]
  idx = {h: i for i, h in enumerate(hdr)}

  def to_bytes(r, name):
    v = float(r[idx[name]].replace(',', ''))
    u = units[idx[name]].lower()
    return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
  res = []
  for r in rows[2:]:
    res.append(dict(kernel=r[idx['Kernel Name']], read=to_bytes(r, 'dram__bytes_read.sum'),
                    write=to_bytes(r, 'dram__bytes_write.sum'),
                    ns=float(r[idx['gpu__time_duration.sum']].replace(',', '')) *
                    {'ns': 1, 'nsecond': 1, 'us': 1e3, 'usecond': 1e3, 'ms': 1e6, 'msecond': 1e6}.get(units[idx['gpu__time_duration.sum']].lower(), 1)))
  return res


def main(src, dst):
  out = {}

  def avg(name, inst=''):
    ls = launches(os.path.join(src, 'r02_%s.ncu-rep' % name))
    # the dim-16 instance is the one the roofline names (the dim-1 wide table runs the same kernels)
    sel = [l for l in ls if inst in l['kernel']] or ls
    return (sum(l['read'] for l in sel) / len(sel), sum(l['write'] for l in sel) / len(sel),
            sum(l['ns'] for l in sel) / len(sel), len(sel))
  fr, fw, fns, fn = avg('fwd_single_kernel', '<4, 4>')
  out['er_embedding_fwd'] = dict(dram_bytes=fr + fw, dram_read=fr, dram_write=fw, launches_averaged=fn,
                                 ns_under_ncu=fns, source='r02_ncu_fwd_single.txt')
  br, bw, bns, bn = avg('bk_fused_kernel', '<4>')
  pr, pw, pns, pn = avg('bk_place_kernel')
  out['er_embedding_bwd'] = dict(dram_bytes=br + bw + pr + pw, dram_read=br + pr, dram_write=bw + pw,
                                 parts={'bk_fused_kernel': dict(read=br, write=bw, ns_under_ncu=bns, launches_averaged=bn),
                                        'bk_place_kernel': dict(read=pr, write=pw, ns_under_ncu=pns, launches_averaged=pn)},
                                 source='r02_ncu_bk_fused.txt')
  json.dump(out, open(dst, 'w'), indent=1)
  print(json.dumps(out, indent=1))


if __name__ == '__main__':
  main(sys.argv[1], sys.argv[2])
