#!/bin/bash
# Is hardware video decode (NVDEC) reachable from this image on the GPU box?
echo "== libnvcuvid / libnvidia-encode in the loader cache"; ldconfig -p | grep -i -E "nvcuvid|nvidia-encode" || echo "none"
echo "== files"; ls -l /usr/lib/x86_64-linux-gnu/libnvcuvid* /usr/lib64/libnvcuvid* /usr/local/nvidia/lib64/libnvcuvid* 2>/dev/null || echo "none"
echo "== headers (nvcuvid.h / cuviddec.h / dynlink_nvcuvid.h)"; find / -xdev \( -name "nvcuvid.h" -o -name "cuviddec.h" -o -name "dynlink_nvcuvid.h" \) 2>/dev/null | head -5; echo "(end)"
echo "== NVIDIA_DRIVER_CAPABILITIES=$NVIDIA_DRIVER_CAPABILITIES"
echo "== cv2.cudacodec"; python -c "import cv2; print(hasattr(cv2, 'cudacodec'), cv2.cuda.getCudaEnabledDeviceCount() if hasattr(cv2,'cuda') else 'no cv2.cuda')" 2>&1 | tail -1
echo "== ffmpeg binary"; which ffmpeg || echo "none"
echo "== PyNvVideoCodec / torchcodec / torchvision.io / decord / av"; for m in PyNvVideoCodec torchcodec torchvision decord av nvidia.dali; do python -c "import $m" 2>/dev/null && echo "$m: importable" || echo "$m: absent"; done
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader
