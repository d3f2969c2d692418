"""Import the UNMODIFIED reference from /root/reference (build container only).

Test infrastructure.  The dataset-only dependencies h5py / skimage /
pycocotools are absent here and are stubbed in sys.modules;
nothing on the hot path touches them.  Returns None when the reference tree
is not present (e.g. on the GPU box)."""
import os
import sys
import types

REF = '/root/reference'


def have_reference():
  return os.path.isdir(os.path.join(REF, 'sg2im'))


def import_reference():
  if not have_reference():
    return None
  for m in ('h5py', 'skimage', 'skimage.transform', 'pycocotools',
            'pycocotools.mask', 'imageio'):
    if m not in sys.modules:
      try:
        __import__(m)
      except Exception:
        mod = types.ModuleType(m)
        sys.modules[m] = mod
        if '.' in m:
          setattr(sys.modules[m.split('.')[0]], m.split('.')[1], mod)
  st = sys.modules['skimage.transform']
  if not hasattr(st, 'resize'):
    st.resize = None                      # only used by the COCO dataset class
  if REF not in sys.path:
    sys.path.insert(0, REF)
  import sg2im.model, sg2im.discriminators, sg2im.losses, sg2im.layout  # noqa
  import sg2im.graph, sg2im.crn, sg2im.bilinear, sg2im.layers  # noqa
  import sg2im
  return sg2im
