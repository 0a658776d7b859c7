import subprocess, tarfile, io, os, sys, re, time
from concurrent.futures import ThreadPoolExecutor
ref='/root/reference'
res=sorted(os.listdir(f'{ref}/test/results'))
jobs=[]
for f in res:
    if not f.endswith('.out.txt.tar.gz') or '-lex' in f: continue
    name=f[:-len('.out.txt.tar.gz')]
    m=re.match(r'(.*)-elim(\d+)$',name)
    ex,el=(m.group(1),int(m.group(2))) if m else (name,0)
    for ms in (None,0):
        jobs.append((f,ex,el,ms))
def run(j):
    f,ex,el,ms=j
    gold=tarfile.open(f'{ref}/test/results/{f}').extractfile(tarfile.open(f'{ref}/test/results/{f}').getmembers()[0]).read()
    out=f'/tmp/gbt/out_{ex}_{el}_{ms}.txt'
    cmd=['./oracle/_build/gamba_oracle','-i',f'{ref}/examples/{ex}.txt','-e',str(el),'-o',out,'-v','0']
    if ms is not None: cmd.append(f'--max-spairs={ms}')
    t=time.time()
    try:
        r=subprocess.run(cmd,capture_output=True,timeout=1200)
    except subprocess.TimeoutExpired:
        return (j,'TIMEOUT',time.time()-t)
    if r.returncode!=0: return (j,'RC%d %s'%(r.returncode,r.stderr[-200:]),time.time()-t)
    ok=open(out,'rb').read()==gold
    os.remove(out)
    return (j,'OK' if ok else 'DIFF',time.time()-t)
with ThreadPoolExecutor(8) as ex:
    bad=0
    for j,st,t in ex.map(run,jobs):
        if st!='OK' or t>20: print(j,st,'%.1fs'%t,flush=True)
        bad+= st!='OK'
print('total',len(jobs),'bad',bad)
