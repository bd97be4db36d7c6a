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

  @doc """
  One handle = one GPU workspace.  A handle is single-flight: keep it in the state of ONE process (one camera, one consumer).
  If the term does reach a second process, a call that finds the handle busy returns `{:error, :busy}`.
  """
  def create(_nfeatures, _scale_factor, _nlevels, _ini_th, _min_th, _device, _max_width, _max_height),
    do: :erlang.nif_error(:nif_not_loaded)

  @doc "as create/8 with room for `max_batch` frames per `extract_batch/5` call"
  def create(_nfeatures, _scale_factor, _nlevels, _ini_th, _min_th, _device, _max_width, _max_height, _max_batch),
    do: :erlang.nif_error(:nif_not_loaded)

  @doc """
  `batch` gray frames back to back in one binary -> `{:ok, counts, mono_indices, keypoints, descriptors, cap}`; frame i owns
  records `i * cap .. i * cap + counts[i] - 1` of the two result binaries (counts / mono_indices: int32 little-endian).
  """
  def extract_batch(_handle, _frames_binary, _batch, _width, _height), do: :erlang.nif_error(:nif_not_loaded)

  @doc "a row shard of a descriptor database resident on the GPU: rows x 32 bytes, `row_offset` = global index of its first row"
  def knn2_create(_descriptors_binary, _device, _row_offset), do: :erlang.nif_error(:nif_not_loaded)

  @doc "k = 2 nearest rows per query (cv::BFMatcher NORM_HAMMING): `{:ok, indices (nq x 2 int32), distances (nq x 2 int32)}`; backend 0 = POPC, 1 = tensor cores (int8), 2 = tensor cores (FP4, fastest); identical results"
  def knn2(_db, _queries_binary, _backend), do: :erlang.nif_error(:nif_not_loaded)

  def extract(_handle, _gray_binary, _width, _height), do: :erlang.nif_error(:nif_not_loaded)

  @doc "format: 1 = RGB, 2 = BGR (an `Evision.Mat` from `Evision.VideoCapture.read/1`), 3 = RGBA, 4 = BGRA; gray conversion on the GPU"
  def extract_color(_handle, _pixels_binary, _width, _height, _format), do: :erlang.nif_error(:nif_not_loaded)

  @doc "the wire's PPM binary (`Evision.imencode(\".ppm\", mat)`, slam_handler.ex:275-277) as it is; camera_rgb = 1 for `rgb: 1`"
  def extract_ppm(_handle, _ppm_binary, _camera_rgb), do: :erlang.nif_error(:nif_not_loaded)

  def match_windowed(_handle, _q_desc, _q_uvr, _q_levels, _t_kp, _t_desc, _bounds),
    do: :erlang.nif_error(:nif_not_loaded)
end
