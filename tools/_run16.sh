python tools/ws_timeline.py run 1024 2>&1 | tail -24
bash tools/_run15.sh
