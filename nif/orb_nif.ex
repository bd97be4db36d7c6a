defmodule SendSlam.OrbNif do
  @moduledoc """
  In-VM ORB extraction / windowed matching on a B200 through liborbx.so (NIF in nif/orbx_nif.c).

  Meant for a consumer process registered in `SendSlam.CameraRegistry` the way `SendSlam.ImageTimer` is
  (send_slam/lib/send_slam/timer.ex:24-27): on `{:camera_frame, {:ok, opts}}` convert `opts[:frame]` to a gray
  `Evision.Mat`, call `extract/4` with `Evision.Mat.to_binary/1`, and ship ~75 KB of keypoints + descriptors instead
  of a 3 MB PPM (send_slam/lib/send_slam/slam_handler.ex:140-156).  All functions return `{:ok, ...}` or
  `{:error, reason_atom}`; they run on dirty IO schedulers.
  """
  @on_load :load_nif
  def load_nif do
    path = :filename.join(:code.priv_dir(:send_slam), ~c"orbx_nif")
    :erlang.load_nif(path, 0)
  end

  def create(_nfeatures, _scale_factor, _nlevels, _ini_th, _min_th, _device, _max_width, _max_height),
    do: :erlang.nif_error(:nif_not_loaded)

  def extract(_handle, _gray_binary, _width, _height), do: :erlang.nif_error(:nif_not_loaded)

  @doc "format: 1 = RGB, 2 = BGR (an `Evision.Mat` from `Evision.VideoCapture.read/1`), 3 = RGBA, 4 = BGRA; gray conversion on the GPU"
  def extract_color(_handle, _pixels_binary, _width, _height, _format), do: :erlang.nif_error(:nif_not_loaded)

  @doc "the wire's PPM binary (`Evision.imencode(\".ppm\", mat)`, slam_handler.ex:275-277) as it is; camera_rgb = 1 for `rgb: 1`"
  def extract_ppm(_handle, _ppm_binary, _camera_rgb), do: :erlang.nif_error(:nif_not_loaded)

  def match_windowed(_handle, _q_desc, _q_uvr, _q_levels, _t_kp, _t_desc, _bounds),
    do: :erlang.nif_error(:nif_not_loaded)
end
