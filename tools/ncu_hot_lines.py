import csv, subprocess, sys, io
rep=sys.argv[1]; n=int(sys.argv[2]) if len(sys.argv)>2 else 40
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
cur=None; hdr=None; out=[]
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r and r[0]=='Line No': si=r.index('# Samples'); ii=r.index('Instructions Executed'); hdr=1; continue
    if cur and hdr and r and r[0].isdigit():
        try: out.append((int(r[si]), int(r[ii]), cur, r[0], r[1].strip()[:110]))
        except: pass
tot=sum(o[0] for o in out); toti=sum(o[1] for o in out)
print('samples',tot,'instr',toti)
for s,ni,f,ln,t in sorted(out,reverse=True)[:n]: print(f'{100*s/tot:5.1f}% {100*ni/toti:5.1f}%i {f}:{ln} {t}')
