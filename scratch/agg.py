import csv, re, collections, sys
fn=sys.argv[1]
with open(fn) as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
seq=[]
for row in r:
    name=row['Kernel Name']; name=re.sub(r'\(.*','',name); name=re.sub(r'^void ','',name).replace('<unnamed>::','')
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    if u=='ns': v/=1000
    seq.append((name,v,row.get('Grid Size','')))
idx=[i for i,(n,_,_) in enumerate(seq) if n.startswith('k_zero_flags')]
a=idx[-2]; b=idx[-1]
step=[x for x in seq[a:b] if not x[0].startswith(('k_spec','k_patch','k_add_bn','k_ln_affine','at::','k_post','k_unpack','k_pack'))]
# drop specformer gemms: they come after k_step_inc; cut at k_step_inc
cut=[i for i,x in enumerate(step) if x[0].startswith('k_step_inc')]
if cut: step=step[:cut[0]+1]
tot=sum(v for _,v,_ in step)
print('step launches',len(step),'total us',tot)
agg=collections.defaultdict(lambda:[0,0.0])
for n,v,_ in step: agg[n][0]+=1; agg[n][1]+=v
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print('%-60s n=%3d total %8.1f avg %7.1f  %5.1f%%'%(k[:60],v[0],v[1],v[1]/v[0],100*v[1]/tot))
if len(sys.argv)>2:
    start=[i for i,x in enumerate(step) if x[0].startswith('k_rbf')][3]
    for n,v,g in step[start:start+24]: print('%-55s %8.1f us grid %s'%(n[:55],v,g))
