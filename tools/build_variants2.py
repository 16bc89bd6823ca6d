import subprocess,sys,os
sys.path.insert(0,'/root/repo')
# name:TPB:MINB:T:STAGES:OSTAGES
import acme_jl_b200._build as b
b.build()
libs='/root/repo/tools/libs'
procs=[]
for spec in sys.argv[1:]:
    name,tpb,mb,T,st,ost=spec.split(':')
    obj=f'{libs}/tpi_{name}.o'
    cmd=[b.nvcc()]+b.NVCC_FLAGS+[f'-DACME_TPI_TPB={tpb}',f'-DACME_TPI_MINB={mb}',f'-DACME_TPI_T={T}',f'-DACME_TPI_STAGES={st}',f'-DACME_TPI_OSTAGES={ost}','-c','-o',obj,os.path.join(b.CSRC,'tpi.cu')]
    procs.append((name,obj,subprocess.Popen(cmd,stdout=subprocess.PIPE,stderr=subprocess.STDOUT,text=True)))
for name,obj,p in procs:
    out=p.communicate()[0]
    if p.returncode: print(name,'FAILED',out[-1500:]); continue
    lib=f'{libs}/lib_{name}.so'
    objs=[obj if s=='tpi.cu' else b._obj(s) for s in b.SOURCES]
    r=subprocess.run([b.nvcc(),'-shared','-cudart','static','-gencode','arch=compute_100a,code=sm_100a','-Xcompiler','-fPIC','-o',lib]+objs,capture_output=True,text=True)
    print(name,'->',lib if r.returncode==0 else r.stderr[-300:]); os.remove(obj)
