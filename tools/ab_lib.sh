#!/bin/bash
# A/B of library builds ON THE GPU BOX: three interleaved bench runs per library (value, sustained, model-only img/s, roofline.frac).
#   make -C multiyolov5_b200/csrc SFU_EVERY=3 BUILD=build_sfu3 OUT=../libmyolo_sfu3.so ; bash tools/ab_lib.sh libmyolo_sm100a libmyolo_sfu3
LIBS=${@:-libmyolo_sm100a}
for rep in 1 2 3; do for lib in $LIBS; do
  MYOLO_LIB=multiyolov5_b200/$lib.so python bench.py --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['value']), round(d['sustained']['value']), round(d['model_only_images_per_s']), round(d['roofline']['frac'],4))"
done; done
