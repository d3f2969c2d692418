"""CPU oracle of the accelerated path — TEST INFRASTRUCTURE ONLY.

Importable from ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
legs; never from ``sg2im_b200`` (tests/test_host_cpu.py checks that the product
package has no import of it).
  sg2im_oracle       generator + discriminators + training iteration
  validation_oracle  de-normalisation, box IoU, the validation pass
  relations_oracle   COCO scene-graph synthesis of one sample
"""
