import csv,sys,subprocess
rep=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 40
out=subprocess.run(['ncu','-i',rep,'--page','source','--print-source','cuda,sass','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
h=None
for i,r in enumerate(rows):
    if 'Instructions Executed' in r: h=r; start=i+1; break
ci=h.index('Instructions Executed'); cs=h.index('Source'); csamp=h.index('# Samples')
# rows: cuda lines have line numbers in col0; sass rows have addresses. aggregate by cuda line rows
data=[]
for r in rows[start:]:
    if len(r)<=ci: continue
    try: n=int(r[ci] or 0)
    except: continue
    try: sm=int(r[csamp] or 0)
    except: sm=0
    data.append((n,sm,r[0],r[cs]))
cuda=[d for d in data if not d[2].startswith('0x') and d[2].strip().isdigit()]
tot=sum(d[0] for d in cuda); tots=sum(d[1] for d in cuda)
print('total warp inst (cuda lines)',tot,'samples',tots)
for d in sorted(cuda,reverse=True)[:top]:
    print(f"{d[0]/tot*100:5.1f}% inst  {d[1]/max(1,tots)*100:5.1f}% samp  L{d[2]:>4}: {d[3].strip()[:120]}")
