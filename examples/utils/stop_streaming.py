"""Ask a running streaming job to stop: sends STOP to its reservation server
(reference: examples/utils/stop_streaming.py:9-18).   usage: stop_streaming.py <host> <port>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

if __name__ == "__main__":
  from tensorflowonspark_b200 import reservation
  host, port = sys.argv[1], int(sys.argv[2])
  client = reservation.Client((host, port))
  client.request_stop()
  client.close()
  print("stop requested at {}:{}".format(host, port))
