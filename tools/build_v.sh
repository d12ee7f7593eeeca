#!/bin/bash
# build with ptxas -v and print one line per step kernel: name, registers, spills
cd "$(dirname "$0")/../pybitup_b200/csrc" || exit 1
make -j8 NVCCFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v" > /tmp/build_v.log 2>&1
rc=$?
grep -E "error" /tmp/build_v.log | sort | uniq -c | head -20
python3 - <<'PY'
import re,subprocess
txt=open('/tmp/build_v.log').read()
pat=re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\s*\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\s*\n.*?Used (\d+) registers",re.S)
names=[m for m in pat.findall(txt)]
if names:
    dem=subprocess.run(['c++filt']+[n[0] for n in names],capture_output=True,text=True).stdout.splitlines()
    for d,n in zip(dem,names):
        if 'mh_step' in d or 'isde' in d or 'propag' in d:
            print("%-75s regs=%s stack=%s spill_st=%s spill_ld=%s"%(d.replace('pbu::','').replace('void ','')[:75],n[4],n[1],n[2],n[3]))
PY
exit $rc
