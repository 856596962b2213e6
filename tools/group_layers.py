import json, collections, sys
d=json.load(open(sys.argv[1]))
g=collections.defaultdict(lambda:[0,0.0,0.0])
for r in d['layers']:
    if r['name']=='spp': print('spp %.3f'%r['ms']); continue
    key=(r['hw'], r['k'], r['s'], r['cin'], r['cout'], r['bn'], r.get('mode',0), r.get('epi',0), r.get('st',0), r.get('grp',0), r.get('cps',0), r.get('nepi',0), r.get('bres',0))
    g[key][0]+=1; g[key][1]+=r['ms']; g[key][2]+=r['gflop']
print('total %.3f ms -> %.0f img/s'%(d['total_ms'], d['img_per_s']))
print('(hw,k,s,cin,cout,bn,mode,epi,stages,group,ctas/sm,epi warps,resident W) n  ms  TF/s | hbm-min ms | mma-min ms(1423TF)')
for k,v in sorted(g.items(), key=lambda kv:-kv[1][1]):
    hw,kk,s,cin,cout,bn=k[:6]
    rows=d['batch']*(hw+2)**2; inrows=d['batch']*((hw*s)+2)**2
    byt=(inrows*cin+rows*cout)*2*v[0]
    print(k, v[0], '%.3f'%v[1], '%.0f'%(v[2]/v[1]), '| %.3f'%(byt/6.5e9), '| %.3f'%(v[2]/1423))
