for cfg in "XR_DUAL_PINS=8 XR_DUAL_MINC=8" "XR_DUAL_PINS=4 XR_DUAL_MINC=8" "XR_DUAL_PINS=4 XR_DUAL_MINC=4" "XR_DUAL_PINS=6 XR_DUAL_MINC=4" "XR_DUAL_PINS=3 XR_DUAL_MINC=4" "XR_DUAL_PINS=8 XR_DUAL_MINC=8 XR_MEDIUM_PINS=3" "XR_DUAL_PINS=8 XR_DUAL_MINC=16"; do
echo "== $cfg"; env $cfg python tools/diag_route.py 2>&1 | grep -o "ms/step mean [0-9.]* median [0-9.]*\|timeline.*" ; done
