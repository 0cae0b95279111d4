set +e
echo "== nodes"; ls /sys/devices/system/node/ | tr '\n' ' '; echo
for n in /sys/devices/system/node/node*; do echo "$n cpus=$(cat $n/cpulist) mem=$(grep MemTotal $n/meminfo | awk '{print $4,$5}')"; done
echo "== self"; grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
echo "== cgroup"; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null
echo "== gpus"; nvidia-smi --query-gpu=index,pci.bus_id --format=csv,noheader | while IFS=, read i b; do b=$(echo $b | tr 'A-Z' 'a-z' | sed 's/^ *0000//'); echo "gpu $i $b numa=$(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null) local_cpus=$(cat /sys/bus/pci/devices/$b/local_cpulist 2>/dev/null)"; done
echo "== topo"; nvidia-smi topo -m 2>&1 | head -24
echo "== lscpu"; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" 
python - <<'P'
import ctypes, os
libc = ctypes.CDLL("libc.so.6", use_errno=True)
# get_mempolicy syscall 239, set_mempolicy 238, mbind 237 on x86_64
mode = ctypes.c_int(); mask = (ctypes.c_ulong * 16)()
r = libc.syscall(239, ctypes.byref(mode), mask, 1024, 0, 0)
print("get_mempolicy rc", r, "mode", mode.value, "errno", ctypes.get_errno())
for node in (0, 1):
    m = (ctypes.c_ulong * 16)(); m[0] = 1 << node
    r = libc.syscall(238, 2, m, 1024)   # MPOL_BIND
    print("set_mempolicy(BIND, node", node, ") rc", r, "errno", ctypes.get_errno())
libc.syscall(238, 0, None, 0)
P
