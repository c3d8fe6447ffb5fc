"""Host-side placement for the host-buffer entry points (se3ds_reproject_host, guidance.HostReprojector).

A multi-socket host reaches a GPU fastest from the memory of the socket its PCIe root port hangs off.
Pinned memory is placed by first touch, so a process that feeds one GPU binds itself to that socket's
cores before it allocates its pinned buffers.  Pure sysfs + sched_setaffinity: no libnuma needed."""
import os
from typing import Optional

import torch


def _read(path: str) -> Optional[str]:
  try:
    with open(path) as f:
      return f.read().strip()
  except OSError:
    return None


def _parse_cpulist(text: str):
  cpus = set()
  for part in text.split(','):
    if not part:
      continue
    lo, _, hi = part.partition('-')
    cpus.update(range(int(lo), int(hi or lo) + 1))
  return cpus


def gpu_numa_node(device: int = 0) -> Optional[int]:
  """NUMA node of the GPU's PCIe function, or None when the platform does not say (VMs often report -1)."""
  props = torch.cuda.get_device_properties(device)
  bdf = '%04x:%02x:%02x.0' % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
  text = _read('/sys/bus/pci/devices/%s/numa_node' % bdf)
  if text is None:
    return None
  node = int(text)
  return node if node >= 0 else None


def bind_to_gpu_numa_node(device: int = 0) -> dict:
  """Restricts this process to the cores of the GPU's NUMA node (so first-touch places pinned buffers there).
  -> {'node': n or None, 'cpus': count bound or None, 'nodes_online': ...}; a no-op when the node is unknown,
  has no cores allowed to this process, or the host has a single node."""
  info = {'node': gpu_numa_node(device), 'cpus': None, 'nodes_online': _read('/sys/devices/system/node/online')}
  if info['node'] is None or info['nodes_online'] in (None, '0'):
    return info
  text = _read('/sys/devices/system/node/node%d/cpulist' % info['node'])
  if not text:
    return info
  cpus = _parse_cpulist(text) & os.sched_getaffinity(0)
  if cpus:
    os.sched_setaffinity(0, cpus)
    info['cpus'] = len(cpus)
  return info
