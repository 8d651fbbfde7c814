#!/bin/bash
# vpt_offline end to end on the GPU box: settings + scene files written here (the reference tree is not on the box), 4 frames.
set -e
D=gpurun_out/offline_demo; mkdir -p $D/data/settings $D/data/scene
cat > $D/data/settings/global_settings.yaml <<Y
denoising:
  atrousIterationNum: 1
postprocess:
  manualExposure: 0.8
  toneMappingCurve: 0
sky:
  timeOfDay: 0.25
  sunAxisAngle: 45
  sunAxisRotate: 0
  skyBrightness: 1
Y
cat > $D/data/scene/scene_export.yaml <<Y
camera:
  position: [35.6184, 11.8733, 42.0387]
  direction: [-0.321564, -0.0129988, -0.946799]
  up: [0, 1, 0]
  fov: 90
Y
P=real-time-path-tracing-voxel-blocks_b200
$P/vpt_offline --width 640 --height 360 --frames 4 --spp 4 --output $D/offline_render --scene $D/data/scene/scene_export.yaml \
  --settings $D/data/settings/global_settings.yaml --tables $P/data/bluenoise_tables.bin --sky-tables $P/data/sky_tables.bin 2>&1 | tail -8
ls -la $D/*.png | head
