import os, glob
print('affinity', sorted(os.sched_getaffinity(0)))
for n in sorted(glob.glob('/sys/devices/system/node/node*')):
    try: print(n, open(n + '/cpulist').read().strip(), open(n + '/meminfo').read().split('\n')[0])
    except Exception as e: print(n, e)
try:
    import pynvml
    pynvml.nvmlInit()
    for i in range(pynvml.nvmlDeviceGetCount()):
        h = pynvml.nvmlDeviceGetHandleByIndex(i)
        pci = pynvml.nvmlDeviceGetPciInfo(h)
        bus = pci.busId.decode() if isinstance(pci.busId, bytes) else pci.busId
        try: aff = pynvml.nvmlDeviceGetCpuAffinity(h, 8)
        except Exception as e: aff = e
        p = '/sys/bus/pci/devices/' + bus.lower()[-12:] + '/numa_node'
        try: node = open(p).read().strip()
        except Exception as e: node = repr(e)
        print('gpu', i, bus, 'numa', node, 'cpu affinity mask', [hex(a) for a in aff] if isinstance(aff, list) else aff)
except Exception as e:
    print('nvml', e)
os.system('nvidia-smi topo -m 2>&1 | head -20')
