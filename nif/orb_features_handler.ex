defmodule SendSlam.OrbFeaturesHandler do
  @moduledoc """
  SURVEY.md §8f-3 — what `SendSlam.SlamHandler.handle_info({:camera_frame, ...})` becomes when extraction runs in the VM
  (send_slam/lib/send_slam/slam_handler.ex:59-88).  It is the reference's own flow with one step exchanged: instead of

      {:ok, ppm}    <- encode_to_ppm(mat),                   # slam_handler.ex:66, 275-277
      {:ok, packet} <- build_frame_packet(ppm, dims, opts)   # :67, 140-156  (2.7 MB at 1280x720)

  the frame goes through `SendSlam.OrbNif` (nif/orbx_nif.c -> liborbx.so on the B200) and the packet carries
  keypoints + descriptors (~75 KB at nFeatures 1250):

      %{type: "features", camera_id, timestamp, width, height, mono_index, n,
        keypoints: <<n * 28 bytes, cv::KeyPoint layout>>, descriptors: <<n * 32 bytes>>}

  Framing is unchanged (`<<len::32-big-unsigned>>` + MessagePack, slam_handler.ex:283-291).  The message layout is the
  one `orbx_wire_pack_features` / `orbx_wire_parse_features` implement (include/orbx_wire.h); the C++ backend needs one
  more branch beside `packet.type == "frame"` (slam_backends/orb_slam_3/orbslam3_mono_networked.cc:520) that calls
  `orbx_wire_parse_features` and an ORB-SLAM3 `Frame` constructor taking precomputed features (INTEGRATION.md §2).

  NOT RUN in this repository: there is no BEAM in the build image.  The C side of every call made here is exercised on a
  B200 through the mock host (tests/test_nif_mock_gpu.py) and the message layout through the Python msgpack package
  (tests/test_abi_cpu.py, tests/test_gpu_parity.py).
  """
  require Logger

  @nfeatures 1250
  @camera_rgb 1

  @doc "Called once per connection, e.g. from handle_connection/2 next to the registry registrations (slam_handler.ex:21-24)."
  def open(max_width \\ 1280, max_height \\ 800, device \\ 0) do
    # the extractor parameters the backend reads from its YAML literal (orbslam3_mono_networked.cc:193-206)
    SendSlam.OrbNif.create(@nfeatures, 1.2, 8, 20, 7, device, max_width, max_height)
  end

  @doc "Drop-in for the `with` chain of slam_handler.ex:63-67: returns `{:ok, iodata_packet}` ready for ThousandIsland.Socket.send/2."
  def build_features_packet(orb, mat, opts) do
    camera_id = Keyword.get(opts, :camera_id, 1)
    timestamp = Keyword.get(opts, :timestamp, System.monotonic_time(:nanosecond) / 1_000_000_000)

    # the same bytes the reference would have put on the wire; the NIF parses the header and converts on the GPU
    ppm = Evision.imencode(".ppm", mat)

    case SendSlam.OrbNif.extract_ppm(orb, ppm, @camera_rgb) do
      {:ok, n, mono_index, keypoints, descriptors, width, height} ->
        payload = %{
          type: "features",
          camera_id: camera_id,
          timestamp: timestamp,
          width: width,
          height: height,
          mono_index: mono_index,
          n: n,
          keypoints: Msgpax.Bin.new(keypoints),
          descriptors: Msgpax.Bin.new(descriptors)
        }

        packed = Msgpax.pack!(payload)
        {:ok, [<<IO.iodata_length(packed)::32-big-unsigned>>, packed]}

      {:error, :empty_image} ->
        # the backend would have logged "Failed to decode frame image data." and skipped the frame (:547-551)
        {:error, :undecodable_frame}

      {:error, reason} ->
        Logger.warning("OrbFeaturesHandler: extraction failed: #{inspect(reason)}")
        {:error, reason}
    end
  end
end
