import sys, collections
rows=[l.split() for l in open(sys.argv[1])]
ev=[]
for r in rows:
    s,e,d=float(r[0]),float(r[1]),float(r[2]); name=' '.join(r[4:])
    kind='H2D' if 'HtoD' in name else 'D2H' if 'DtoH' in name else 'scan' if 'scan_kernel' in name else 'rebuild' if 'rebuild' in name else 'encode' if 'encode_kernel' in name else 'other'
    ev.append((s,e,d,kind,r[3],name.split()[-1] if 'Memcpy' in name else ''))
end=max(e[1] for e in ev)
print('end',end)
for k in ('H2D','D2H','scan','rebuild','encode'):
    tot=sum(e[2] for e in ev if e[3]==k); first=min(e[0] for e in ev if e[3]==k); last=max(e[1] for e in ev if e[3]==k)
    print(k,'busy %.1f first %.1f last %.1f n %d'%(tot,first,last,sum(1 for e in ev if e[3]==k)))
for k in ('H2D','D2H'):
    bins=collections.Counter()
    for e in ev:
        if e[3]==k and e[5].isdigit():
            b=int(e[5]); s0,e0=e[0],e[1]; t=s0
            while t<e0:
                nb=min(e0,(int(t//5)+1)*5)
                bins[int(t//5)]+=b*(nb-t)/max(e0-s0,1e-9); t=nb
    print(k,' '.join('%d:%.0f'%(i,bins[i]/5e6) for i in range(int(end//5)+1)))
if len(sys.argv)>2:
  for e in ev:
    if e[3] in sys.argv[2].split(',') and e[2]>float(sys.argv[3]): print('%.1f %.1f %.2f %s %s %s'%(e[0],e[1],e[2],e[3],e[4],e[5]))
